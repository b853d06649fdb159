// Audio front end at the entry of the path: Kaldi-compatible log-mel filterbank features and
// utterance CMVN on the GPU.  In the reference these run on the CPU inside the data loader
// (ofasys/preprocessor/default/audio.py:283-305 -> _get_torchaudio_fbank :507-516 ->
// torchaudio.compliance.kaldi.fbank(waveform, num_mel_bins=80, sample_frequency=16000), torchaudio 2.11; then
// ofasys/utils/audio_feature_transforms/utterance_cmvn.py:33-44).
//
// One CTA per frame: frame -> remove DC -> pre-emphasis -> window -> zero-pad -> radix-2 FFT in shared memory ->
// power spectrum -> mel filterbank -> log.  HBM traffic: the waveform once (frames overlap 2.5x, served by L2) and
// n_mel floats per frame out; the kernel is latency / shared-memory bound and tiny next to the model step.
#include "common.cuh"

namespace {

constexpr int kFbThreads = 256;

__device__ __forceinline__ float block_sum_fb(float v, float* red /* [8] */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kFbThreads / 32; ++w) t += red[w];
  return t;
}

template <int NFFT, int LOG2N>
__global__ void __launch_bounds__(kFbThreads) fbank_kernel(const float* __restrict__ wav, int64_t wav_bs, const int64_t* __restrict__ lengths,
                                                           int64_t n_samples, const float* __restrict__ window, const float* __restrict__ melT,
                                                           int win, int shift, int n_mel, float preemph, float eps, float* __restrict__ out,
                                                           int64_t* __restrict__ n_frames_out, int max_frames) {
  __shared__ float re[NFFT], im[NFFT], tmp[NFFT], tw_re[NFFT / 2], tw_im[NFFT / 2], red[8];
  const int b = blockIdx.y, f = blockIdx.x, tid = threadIdx.x;
  const int64_t len = lengths != nullptr ? lengths[b] : n_samples;
  const int64_t m = len >= win ? 1 + (len - win) / shift : 0;  // snip_edges = True (kaldi.py _get_strided)
  if (f == 0 && tid == 0 && n_frames_out != nullptr) n_frames_out[b] = m < max_frames ? m : max_frames;
  float* orow = out + ((int64_t)b * max_frames + f) * n_mel;
  if (f >= m) {  // padding frame of a shorter utterance
    for (int i = tid; i < n_mel; i += kFbThreads) orow[i] = 0.f;
    return;
  }
  const float* src = wav + (int64_t)b * wav_bs + (int64_t)f * shift;
  // twiddles exp(-2 pi i k / NFFT)
  for (int k = tid; k < NFFT / 2; k += kFbThreads) {
    float s, c;
    sincospif(-2.0f * (float)k / (float)NFFT, &s, &c);
    tw_re[k] = c;
    tw_im[k] = s;
  }
  float s = 0.f;
  for (int i = tid; i < win; i += kFbThreads) {
    const float x = src[i];
    tmp[i] = x;
    s += x;
  }
  const float mean = block_sum_fb(s, red) / (float)win;  // remove_dc_offset (kaldi.py _get_window); also publishes tmp[]
  // y[j] = (x[j] - mean) - preemph * (x[max(j - 1, 0)] - mean), times the window, zero-padded, stored bit-reversed
  for (int j = tid; j < NFFT; j += kFbThreads) {
    float y = 0.f;
    if (j < win) {
      const float xc = tmp[j] - mean, xp = tmp[j > 0 ? j - 1 : 0] - mean;
      y = (xc - preemph * xp) * window[j];
    }
    const int r = (int)(__brev((unsigned)j) >> (32 - LOG2N));
    re[r] = y;
    im[r] = 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int st = 1; st <= LOG2N; ++st) {  // decimation in time
    const int half = 1 << (st - 1), step = NFFT >> st;
    for (int k = tid; k < NFFT / 2; k += kFbThreads) {
      const int j = k & (half - 1), i0 = ((k >> (st - 1)) << st) + j, i1 = i0 + half;
      const float wr = tw_re[j * step], wi = tw_im[j * step];
      const float xr = re[i1], xi = im[i1];
      const float tr = xr * wr - xi * wi, ti = xr * wi + xi * wr;
      const float ur = re[i0], ui = im[i0];
      re[i0] = ur + tr;
      im[i0] = ui + ti;
      re[i1] = ur - tr;
      im[i1] = ui - ti;
    }
    __syncthreads();
  }
  for (int k = tid; k <= NFFT / 2; k += kFbThreads) tmp[k] = re[k] * re[k] + im[k] * im[k];  // use_power
  __syncthreads();
  for (int t = tid; t < n_mel; t += kFbThreads) {
    float acc = 0.f;
    for (int k = 0; k <= NFFT / 2; ++k) acc = fmaf(tmp[k], melT[(int64_t)k * n_mel + t], acc);
    orow[t] = logf(fmaxf(acc, eps));  // use_log_fbank: max(., float32 eps).log()
  }
}

// utterance CMVN in place: one CTA per utterance, one thread per feature column
__global__ void __launch_bounds__(128) cmvn_kernel(float* __restrict__ feats, const int64_t* __restrict__ n_frames, int max_frames, int n_feat,
                                                   int norm_means, int norm_vars) {
  const int b = blockIdx.x;
  const int64_t n = n_frames != nullptr ? n_frames[b] : max_frames;
  if (n <= 0) return;
  float* base = feats + (int64_t)b * max_frames * n_feat;
  for (int c = threadIdx.x; c < n_feat; c += 128) {
    // E[x^2] - mean^2 as the reference computes it, but accumulated in double: the float32 form loses ~2 digits to
    // cancellation (x ~ 10, var ~ 1) -- the reference's own numpy float32 result carries that error
    double s = 0.0, q = 0.0;
    for (int64_t r = 0; r < n; ++r) {
      const double x = (double)base[r * n_feat + c];
      s += x;
      q += x * x;
    }
    const float mean = (float)(s / (double)n);
    const float var = (float)(q / (double)n - (s / (double)n) * (s / (double)n));
    const float inv = norm_vars ? 1.0f / sqrtf(fmaxf(var, 1e-10f)) : 1.0f;
    for (int64_t r = 0; r < n; ++r) {
      float x = base[r * n_feat + c];
      if (norm_means) x -= mean;
      base[r * n_feat + c] = x * inv;
    }
  }
}

}  // namespace

extern "C" int ofab_fbank(const float* wav, int64_t wav_bs, const int64_t* lengths, int B, int64_t n_samples, const float* window,
                          const float* melT, int win, int shift, int nfft, int n_mel, float preemph, float* out, int64_t* n_frames,
                          int max_frames, ofab_stream_t stream) {
  OFAB_REQUIRE(wav != nullptr && window != nullptr && melT != nullptr && out != nullptr, "ofab_fbank: NULL argument");
  OFAB_REQUIRE(B > 0 && max_frames > 0 && n_mel > 0 && win >= 2 && win <= nfft && shift > 0, "ofab_fbank: bad shape B=%d frames=%d n_mel=%d win=%d shift=%d", B, max_frames, n_mel, win, shift);
  OFAB_REQUIRE(nfft == 512 || nfft == 1024 || nfft == 256, "ofab_fbank: nfft=%d (256, 512 or 1024: 8 / 16 / 32 kHz at 25 ms)", nfft);
  OFAB_REQUIRE(B <= 65535, "ofab_fbank: B=%d > 65535", B);
  const float eps = 1.1920928955078125e-07f;  // torch.finfo(float32).eps (kaldi.py _get_epsilon)
  dim3 grid(max_frames, B);
  cudaStream_t st = (cudaStream_t)stream;
#define FB(N, L) fbank_kernel<N, L><<<grid, kFbThreads, 0, st>>>(wav, wav_bs, lengths, n_samples, window, melT, win, shift, n_mel, preemph, eps, out, n_frames, max_frames)
  if (nfft == 512) FB(512, 9);
  else if (nfft == 1024) FB(1024, 10);
  else FB(256, 8);
#undef FB
  OFAB_LAUNCH_CHECK("ofab_fbank");
  return OFAB_OK;
}

extern "C" int ofab_utterance_cmvn(float* feats, const int64_t* n_frames, int B, int max_frames, int n_feat, int norm_means, int norm_vars,
                                   ofab_stream_t stream) {
  OFAB_REQUIRE(feats != nullptr && B > 0 && max_frames > 0 && n_feat > 0, "ofab_utterance_cmvn: bad arguments");
  cmvn_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(feats, n_frames, max_frames, n_feat, norm_means, norm_vars);
  OFAB_LAUNCH_CHECK("ofab_utterance_cmvn");
  return OFAB_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Remaining pieces of the GPU-side preprocessing at the path's entry.
// ------------------------------------------------------------------------------------------------------------------------
namespace {

// SpecAugment masking (utils/audio_feature_transforms/specaugment.py:111-126): for utterance b the host drew n_f frequency
// bands (f0, f) and n_t time bands (t0, t) -- the random draws stay on the host, in the reference's own order -- and every
// element inside a band is overwritten with mask_value (the utterance mean when use_mean, :93-94).
// bands: int32 [B, n_f + n_t, 2]; feats: fp32 [B, max_frames, n_feat] in place; only rows < n_frames[b] are touched.
__global__ void spec_mean_kernel(const float* __restrict__ feats, const int64_t* __restrict__ n_frames, int max_frames, int n_feat, float* __restrict__ mean) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int64_t n = (n_frames ? min((int64_t)max_frames, n_frames[b]) : (int64_t)max_frames) * n_feat;
  const float* x = feats + (int64_t)b * max_frames * n_feat;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) mean[b] = n > 0 ? v / (float)n : 0.f;
  }
}
__global__ void spec_mask_kernel(float* __restrict__ feats, const int64_t* __restrict__ n_frames, int max_frames, int n_feat, const int* __restrict__ bands,
                                 int n_f, int n_t, float mask_value, const float* __restrict__ mean) {
  const int b = blockIdx.y;
  const int64_t nfr = n_frames ? min((int64_t)max_frames, n_frames[b]) : (int64_t)max_frames;
  const int* bd = bands + (int64_t)b * (n_f + n_t) * 2;
  const float mv = mean ? mean[b] : mask_value;
  float* x = feats + (int64_t)b * max_frames * n_feat;
  const int64_t total = nfr * n_feat;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / n_feat), f = (int)(i % n_feat);
    bool hit = false;
    for (int k = 0; k < n_f; ++k) hit = hit || (f >= bd[2 * k] && f < bd[2 * k] + bd[2 * k + 1]);
    for (int k = 0; k < n_t; ++k) hit = hit || (t >= bd[2 * (n_f + k)] && t < bd[2 * (n_f + k)] + bd[2 * (n_f + k) + 1]);
    if (hit) x[i] = mv;
  }
}

// ToTensor + Normalize of the image preprocessor (preprocessor/default/image.py:110-116 -> torchvision): uint8 [B, H, W, 3]
// (decoded, already resized pixels) -> [B, 3, H, W], y = (x / 255 - mean[c]) / std[c], the same fp32 operations in the same
// order (bit-exact against torchvision's div(255), sub_(mean), div_(std)).
template <typename TO>
__global__ void image_normalize_kernel(const uint8_t* __restrict__ img, int64_t HW, int64_t total, float m0, float m1, float m2, float s0, float s1, float s2,
                                       TO* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, p = i % HW;
    const uint8_t* px = img + i * 3;
    TO* o = out + b * 3 * HW + p;
    const float v0 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[0], 255.0f), m0), s0);
    const float v1 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[1], 255.0f), m1), s1);
    const float v2 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[2], 255.0f), m2), s2);
    if (sizeof(TO) == 4) {
      o[0] = (TO)v0; o[HW] = (TO)v1; o[2 * HW] = (TO)v2;
    } else {
      o[0] = (TO)__float2bfloat16(v0); o[HW] = (TO)__float2bfloat16(v1); o[2 * HW] = (TO)__float2bfloat16(v2);
    }
  }
}

// coordinate -> `<bin>_k` token id (preprocessor/default/box.py:101-110): k = round_half_even(x / max_size * (n_bins - 1)),
// id = first_bin_id + k.  fp32 division and multiplication in the reference's order (bit-exact against torch.round).
__global__ void box_bins_kernel(const float* __restrict__ coords, int64_t n, float max_size, int n_bins, int64_t first_bin_id, int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = first_bin_id + (int64_t)rintf(__fmul_rn(__fdiv_rn(coords[i], max_size), (float)(n_bins - 1)));
}

}  // namespace

extern "C" int ofab_spec_augment(float* feats, const int64_t* n_frames, int B, int max_frames, int n_feat, const int32_t* bands, int n_f, int n_t,
                                 float mask_value, int use_mean, float* mean_scratch, ofab_stream_t stream) {
  OFAB_REQUIRE(feats && B > 0 && max_frames > 0 && n_feat > 0 && n_f >= 0 && n_t >= 0, "ofab_spec_augment: bad arguments");
  OFAB_REQUIRE(n_f + n_t == 0 || bands != nullptr, "ofab_spec_augment: bands NULL");
  OFAB_REQUIRE(!use_mean || mean_scratch != nullptr, "ofab_spec_augment: use_mean needs mean_scratch [B]");
  if (n_f + n_t == 0) return OFAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (use_mean) {
    spec_mean_kernel<<<B, 256, 0, st>>>(feats, n_frames, max_frames, n_feat, mean_scratch);
    OFAB_LAUNCH_CHECK("ofab_spec_augment mean");
  }
  const int64_t per = (int64_t)max_frames * n_feat;
  dim3 grid((unsigned)((per + 255) / 256 < 64 ? (per + 255) / 256 : 64), B);
  spec_mask_kernel<<<grid, 256, 0, st>>>(feats, n_frames, max_frames, n_feat, bands, n_f, n_t, mask_value, use_mean ? mean_scratch : nullptr);
  OFAB_LAUNCH_CHECK("ofab_spec_augment");
  return OFAB_OK;
}

extern "C" int ofab_image_normalize(const uint8_t* img, int B, int H, int W, const float* mean3, const float* std3, void* out, int out_dt,
                                    ofab_stream_t stream) {
  OFAB_REQUIRE(img && out && B > 0 && H > 0 && W > 0 && mean3 && std3, "ofab_image_normalize: bad arguments");
  OFAB_REQUIRE(out_dt == OFAB_F32 || out_dt == OFAB_BF16, "ofab_image_normalize: bad out_dt");
  const int64_t HW = (int64_t)H * W, total = HW * B;
  const int64_t want = (total + 255) / 256;
  const int grid = (int)(want < (int64_t)ofab_sm_count() * 8 ? want : (int64_t)ofab_sm_count() * 8);
  if (out_dt == OFAB_F32)
    image_normalize_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(img, HW, total, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], (float*)out);
  else
    image_normalize_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(img, HW, total, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], (bf16*)out);
  OFAB_LAUNCH_CHECK("ofab_image_normalize");
  return OFAB_OK;
}

extern "C" int ofab_box_bins(const float* coords, int64_t n, float max_size, int n_bins, int64_t first_bin_id, int64_t* out, ofab_stream_t stream) {
  OFAB_REQUIRE(coords && out && n > 0 && max_size > 0.f && n_bins > 1, "ofab_box_bins: bad arguments");
  box_bins_kernel<<<(unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0, (cudaStream_t)stream>>>(coords, n, max_size, n_bins, first_bin_id, out);
  OFAB_LAUNCH_CHECK("ofab_box_bins");
  return OFAB_OK;
}
