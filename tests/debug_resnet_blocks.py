"""Debug aid: per-stage error of the channel-last ResNet against the oracle restatement (isolated: every stage is
fed the oracle's input rounded to bf16)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from ofasys_b200.module import resnet as R
from oracle import oracle_model as om

dev = "cuda"
torch.manual_seed(0)
net = R.resnet50_backbone().to(dev)
gen = torch.Generator().manual_seed(1)
with torch.no_grad():
    for n, p in net.named_parameters():
        if p.dim() == 4:
            fan = p.shape[1] * p.shape[2] * p.shape[3]
            p.copy_(torch.randn(p.shape, generator=gen) * (2.0 / fan) ** 0.5)
        elif n.endswith("weight"):
            p.copy_(torch.rand(p.shape, generator=gen) + 0.5)
        else:
            p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
net = net.to(torch.bfloat16).train()
sd = {"r." + k: v.detach().float() for k, v in net.state_dict().items() if v.is_floating_point()}


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().bfloat16()


def nchw(y):
    return y.permute(0, 3, 1, 2).float()


img = torch.randn(4, 3, 64, 64, generator=gen).to(dev)
with torch.no_grad():
    # stem
    c_ref = F.conv2d(img.bfloat16().float(), sd["r.conv1.weight"], stride=2, padding=3)
    cols, Ho, Wo = R.ops.im2col_nchw(img, 7, 2, 3, 152)
    w = F.pad(net.conv1.weight.reshape(64, 147), (0, 5))
    c = R.ops.linear(cols, w, None).view(4, Ho, Wo, 64)
    print("stem conv", rel(nchw(c), c_ref))
    b_ref = F.relu(om._bn(c_ref, sd, "r.bn1", True))
    b = R._bn(net.bn1, nhwc(c_ref), relu=True)
    print("stem bn", rel(nchw(b), b_ref))
    x_ref = F.max_pool2d(b_ref, 3, 2, 1)
    print("maxpool", rel(nchw(R.ops.maxpool3x3s2(nhwc(b_ref))), F.max_pool2d(nhwc(b_ref).permute(0, 3, 1, 2).float(), 3, 2, 1)))
    for li, nb in enumerate([3, 4, 6]):
        for bi in range(nb):
            blk = getattr(net, f"layer{li+1}")[bi]
            p = f"r.layer{li+1}.{bi}"
            stride = 2 if (li > 0 and bi == 0) else 1
            xin = nhwc(x_ref)
            xr = xin.permute(0, 3, 1, 2).float()
            o1r = F.conv2d(xr, sd[p + ".conv1.weight"])
            o1 = R._conv1x1(blk.conv1, xin)
            e = [rel(nchw(o1), o1r)]
            a1r = F.relu(om._bn(o1r, sd, p + ".bn1", True))
            e.append(rel(nchw(R._bn(blk.bn1, nhwc(o1r), relu=True)), a1r))
            o2r = F.conv2d(a1r, sd[p + ".conv2.weight"], stride=stride, padding=1)
            e.append(rel(nchw(R._conv3x3(blk.conv2, nhwc(a1r))), F.conv2d(nhwc(a1r).permute(0,3,1,2).float(), sd[p + ".conv2.weight"], stride=stride, padding=1)))
            a2r = F.relu(om._bn(o2r, sd, p + ".bn2", True))
            e.append(rel(nchw(R._bn(blk.bn2, nhwc(o2r), relu=True)), a2r))
            o3r = F.conv2d(a2r, sd[p + ".conv3.weight"])
            e.append(rel(nchw(R._conv1x1(blk.conv3, nhwc(a2r))), o3r))
            if bi == 0:
                dr = F.conv2d(xr, sd[p + ".downsample.0.weight"], stride=stride)
                e.append(rel(nchw(R._conv1x1(blk.downsample[0], xin)), dr))
                idr = om._bn(dr, sd, p + ".downsample.1", True)
                e.append(rel(nchw(R._bn(blk.downsample[1], nhwc(dr))), idr))
            else:
                idr = xr
            outr = F.relu(om._bn(o3r, sd, p + ".bn3", True) + idr)
            e.append(rel(nchw(R._bn(blk.bn3, nhwc(o3r), residual=nhwc(idr), relu=True)), outr))
            whole = rel(nchw(blk(xin)), om._bottleneck(xr, sd, p, stride, bi == 0, True))
            print(p, tuple(xin.shape), "stages", " ".join(f"{v:.4f}" for v in e), "| block", f"{whole:.4f}")
            x_ref = om._bottleneck(x_ref, sd, p, stride, bi == 0, True)
    y = net(img)
    print("chained", rel(nchw(y), om.resnet_backbone(img, sd, "r", "resnet50", True)))
    print("chained(bf16 img)", rel(nchw(y), om.resnet_backbone(img.bfloat16().float(), sd, "r", "resnet50", True)))
