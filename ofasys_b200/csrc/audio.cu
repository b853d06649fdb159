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
