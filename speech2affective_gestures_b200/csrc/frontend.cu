// Input front-end kernels, sm_100a.
//
//  * s2ag_mfcc_features      utils/common.py:340-349 `get_mfcc_features` = librosa.feature.mfcc(y, sr, n_mfcc)/1000 with its
//                            first/second row differences, the per-chunk CPU step of the long-form synthesis loop
//                            (processor_v2.py:1249-1252).  Two kernels:
//      mfcc_melspec_kernel   one CTA per (frame, clip): reflect-padded hann frame -> 2048-point real FFT as a 1024-point
//                            complex Stockham radix-2 FFT in shared memory -> |X|^2 -> slaney mel filter bank (each
//                            filter reads only its own span of bins) -> 10*log10(max(1e-10, .)).  HBM traffic: the
//                            audio is read once from HBM (4x overlap served by L2), 128 floats per frame written.
//      mfcc_dct_kernel       one CTA per clip: top_db clamp against the clip maximum, DCT-II (ortho), 1/1000, row
//                            differences, [3n-5, frames] written in the layout PoseGenerator.forward takes.
//  * s2ag_expand_inputs      processor_v2.py:606-610: int16 audio * audio_max / 32767 and fp16 -> fp32 MFCC on the device,
//                            so the batch crosses PCIe compressed (2 bytes per sample instead of 4).
#include "common.cuh"
#ifndef S2AG_EMU
#include <cuda_fp16.h>
#endif

namespace {

constexpr int NFFT = 2048;
constexpr int NC = NFFT / 2;      // complex FFT length
constexpr int MEL_THREADS = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(MEL_THREADS)
mfcc_melspec_kernel(const float* __restrict__ audio, long lda, int L, int hop, const float* __restrict__ fb,
                    const int* __restrict__ span, int n_mels, float* __restrict__ logmel, int F) {
  __shared__ float2 buf[2][NC];
  __shared__ float2 tw[NC / 2];    // exp(-2 pi i t / 1024)
  __shared__ float2 tw2[NC / 2 + 1];  // exp(-2 pi i k / 2048), k = 0..512
  const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* x = audio + (long)b * lda;
  for (int t = tid; t < NC / 2; t += MEL_THREADS) {
    float s, c;
    sincospif(-(float)t / (float)(NC / 2), &s, &c);
    tw[t] = make_float2(c, s);
  }
  for (int t = tid; t <= NC / 2; t += MEL_THREADS) {
    float s, c;
    sincospif(-(float)t / (float)NC, &s, &c);
    tw2[t] = make_float2(c, s);
  }
  // frame f covers padded samples [f*hop, f*hop + 2048) of the signal reflect-padded by 1024 on both sides
  for (int j = tid; j < NC; j += MEL_THREADS) {
    float v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int n = 2 * j + u;
      int src = f * hop + n - NFFT / 2;
      if (src < 0) src = -src;
      if (src >= L) src = 2 * (L - 1) - src;
      const float w = 0.5f - 0.5f * cospif((float)n / (float)(NFFT / 2));   // periodic hann
      v[u] = (src >= 0 && src < L) ? x[src] * w : 0.f;
    }
    buf[0][j] = make_float2(v[0], v[1]);
  }
  __syncthreads();
  int cur = 0;
  for (int Ns = 1; Ns < NC; Ns <<= 1) {
    for (int j = tid; j < NC / 2; j += MEL_THREADS) {
      const int k = j & (Ns - 1);
      const float2 a = buf[cur][j];
      const float2 bb = cmul(buf[cur][j + NC / 2], tw[k * ((NC / 2) / Ns)]);
      const int j0 = ((j - k) << 1) + k;
      buf[cur ^ 1][j0] = make_float2(a.x + bb.x, a.y + bb.y);
      buf[cur ^ 1][j0 + Ns] = make_float2(a.x - bb.x, a.y - bb.y);
    }
    cur ^= 1;
    __syncthreads();
  }
  // real-FFT unpack: X[k] = E + e^{-2 pi i k / 2048} O, E = (Z[k] + conj Z[N-k]) / 2, O = -i (Z[k] - conj Z[N-k]) / 2
  float* power = reinterpret_cast<float*>(buf[cur ^ 1]);   // 1025 floats fit in the idle half
  const float2* Z = buf[cur];
  for (int k = tid; k <= NC; k += MEL_THREADS) {
    const float2 zk = Z[k & (NC - 1)];
    const float2 zn = Z[(NC - k) & (NC - 1)];
    const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
    const float2 d = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y + zn.y));
    const float2 o = make_float2(d.y, -d.x);   // -i * d
    // twiddle for k in (512, 1024]: e^{-2 pi i k/2048} = -conj(e^{-2 pi i (1024-k)/2048})
    float2 w = k <= NC / 2 ? tw2[k] : make_float2(-tw2[NC - k].x, tw2[NC - k].y);
    const float2 wo = cmul(w, o);
    const float re = e.x + wo.x, im = e.y + wo.y;
    power[k] = re * re + im * im;
  }
  __syncthreads();
  for (int m = tid; m < n_mels; m += MEL_THREADS) {
    const int lo = span[2 * m], hi = span[2 * m + 1];
    const float* w = fb + (long)m * (NC + 1);
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc += __ldg(w + k) * power[k];
    logmel[((long)b * F + f) * n_mels + m] = 10.f * log10f(fmaxf(acc, 1e-10f));
  }
}

__global__ void __launch_bounds__(256)
mfcc_dct_kernel(const float* __restrict__ logmel, const float* __restrict__ dct, int n_mels, int n_mfcc, int F,
                float top_db, float scale, float* __restrict__ out) {
  S2AG_DYN_SMEM(float, sm);
  float* ls = sm;                       // [F][n_mels]
  float* mf = sm + (long)F * n_mels;    // [n_mfcc][F]
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* src = logmel + (long)b * F * n_mels;
  float mx = -3.0e38f;
  for (int i = tid; i < F * n_mels; i += 256) { const float v = src[i]; ls[i] = v; mx = fmaxf(mx, v); }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  const float floor_db = mx - top_db;
  for (int i = tid; i < n_mfcc * F; i += 256) {
    const int c = i / F, f = i % F;
    const float* d = dct + (long)c * n_mels;
    const float* l = ls + (long)f * n_mels;
    float acc = 0.f;
    for (int m = 0; m < n_mels; ++m) acc += __ldg(d + m) * fmaxf(l[m], floor_db);
    mf[c * F + f] = acc * scale;
  }
  __syncthreads();
  // rows: [0, n) mfcc; [n, 2n-2) d1[i] = mfcc[i+2] - mfcc[i+1]; [2n-2, 3n-5) d2[i] = d1[i+1] - d1[i]
  const int n = n_mfcc, rows = 3 * n - 5;
  float* o = out + (long)b * rows * F;
  for (int i = tid; i < rows * F; i += 256) {
    const int r = i / F, f = i % F;
    float v;
    if (r < n) v = mf[r * F + f];
    else if (r < 2 * n - 2) { const int q = r - n; v = mf[(q + 2) * F + f] - mf[(q + 1) * F + f]; }
    else {
      const int q = r - (2 * n - 2);
      const float a = mf[(q + 3) * F + f] - mf[(q + 2) * F + f];
      const float c = mf[(q + 2) * F + f] - mf[(q + 1) * F + f];
      v = a - c;
    }
    o[i] = v;
  }
}

__global__ void expand_audio_kernel(const short* __restrict__ a16, const void* __restrict__ amax, int max_is_f64,
                                    float* __restrict__ out, long B, long L) {
  const long total = B * L;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long b = e / L;
    float v;
    if (max_is_f64) v = (float)((double)a16[e] * ((const double*)amax)[b] / 32767.0);
    else {
#ifdef S2AG_EMU
      volatile float t = (float)a16[e] * ((const float*)amax)[b];
      v = t / 32767.f;
#else
      v = __fdiv_rn(__fmul_rn((float)a16[e], ((const float*)amax)[b]), 32767.f);   // numpy's two fp32 roundings
#endif
    }
    out[e] = v;
  }
}

__global__ void expand_half_kernel(const unsigned short* __restrict__ h, float* __restrict__ out, long n) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
#ifdef S2AG_EMU
    const unsigned short u = h[e];
    const unsigned sign = (u >> 15) & 1u, ex = (u >> 10) & 31u, man = u & 1023u;
    float v;
    if (ex == 0) v = ldexpf((float)man, -24);
    else if (ex == 31) v = man ? NAN : INFINITY;
    else v = ldexpf((float)(man | 1024u), (int)ex - 25);
    out[e] = sign ? -v : v;
#else
    out[e] = __half2float(__ushort_as_half(h[e]));
#endif
  }
}

}  // namespace

extern "C" int s2ag_mfcc_features(const float* audio, long lda, int B, int L, int hop, const float* mel_fb,
                                  const int* mel_span, int n_mels, const float* dct, int n_mfcc, float top_db,
                                  float scale, float* logmel_ws, float* out, void* stream) {
  S2AG_CHECK_ARG(audio && mel_fb && mel_span && dct && logmel_ws && out);
  S2AG_CHECK_ARG(B >= 0 && L > NFFT / 2 && hop > 0 && lda >= L && n_mels > 0 && n_mels <= 256 && n_mfcc >= 3);
  if (B == 0) return S2AG_OK;
  const int F = 1 + L / hop;
  S2AG_CHECK_ARG((long)F * n_mels * 4 + (long)n_mfcc * F * 4 <= 200 * 1024);
  auto k1 = &mfcc_melspec_kernel;
  S2AG_LAUNCH(k1, dim3(F, B), MEL_THREADS, 0, stream, audio, lda, L, hop, mel_fb, mel_span, n_mels, logmel_ws, F);
  S2AG_CHECK_LAUNCH();
  auto k2 = &mfcc_dct_kernel;
  const size_t smem = ((size_t)F * n_mels + (size_t)n_mfcc * F) * sizeof(float);
#ifndef S2AG_EMU
  if (smem > 48 * 1024) cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  S2AG_LAUNCH(k2, B, 256, smem, stream, logmel_ws, dct, n_mels, n_mfcc, F, top_db, scale, out);
  S2AG_CHECK_LAUNCH();
  return S2AG_OK;
}

extern "C" int s2ag_expand_inputs(const void* audio_i16, const void* audio_max, int max_is_f64, float* audio_out,
                                  long B, long L, const void* mfcc_f16, float* mfcc_out, long n_mfcc, void* stream) {
  S2AG_CHECK_ARG(B >= 0 && L >= 0 && n_mfcc >= 0);
  if (audio_i16 && B * L > 0) {
    S2AG_CHECK_ARG(audio_max && audio_out);
    long blocks = (B * L + 1023) / 1024; if (blocks > 148 * 16) blocks = 148 * 16;
    auto k = &expand_audio_kernel;
    S2AG_LAUNCH(k, (int)blocks, 256, 0, stream, (const short*)audio_i16, audio_max, max_is_f64, audio_out, B, L);
    S2AG_CHECK_LAUNCH();
  }
  if (mfcc_f16 && n_mfcc > 0) {
    S2AG_CHECK_ARG(mfcc_out);
    long blocks = (n_mfcc + 1023) / 1024; if (blocks > 148 * 16) blocks = 148 * 16;
    auto k = &expand_half_kernel;
    S2AG_LAUNCH(k, (int)blocks, 256, 0, stream, (const unsigned short*)mfcc_f16, mfcc_out, n_mfcc);
    S2AG_CHECK_LAUNCH();
  }
  return S2AG_OK;
}
