"""BERT-style initialisation applied by GeneralistModel.initialize (ofasys/module/initialize.py:10-40):
N(0, 0.02) for every nn.Linear / nn.Embedding weight (this also overrides the rel-pos tables'
zero_init) and the q/k/v projections; biases 0; padding row 0."""
import torch.nn as nn


def init_bert_params(module):
    from .multihead_attention import MultiheadAttention

    def normal_(data):
        data.copy_(data.cpu().normal_(mean=0.0, std=0.02).to(data.device))

    if isinstance(module, nn.Linear):
        normal_(module.weight.data)
        if module.bias is not None:
            module.bias.data.zero_()
    if isinstance(module, nn.Embedding):
        normal_(module.weight.data)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()
    if isinstance(module, MultiheadAttention):
        normal_(module.q_proj.weight.data)
        normal_(module.k_proj.weight.data)
        normal_(module.v_proj.weight.data)
