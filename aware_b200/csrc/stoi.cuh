// STOI (short-time objective intelligibility) for a batch of equal-length clip pairs at 10 kHz, as the
// reference computes it through pystoi 0.4.1 (reference metrics/audio.py:43-64 `STOI.__call__`,
// scripts/test.py:86-88; algorithm: Taal et al., IEEE TASL 2011; restated in oracle/stoi_oracle.py --
// pystoi itself is absent here, parity UNPINNED, the CUDA path is tested against that restatement).
//
//   k_stoi_energy   windowed frame energies of the CLEAN signal (frame 256, hop 128) + per-clip maximum
//   k_stoi_scan     silence gate: keep frames within 40 dB of the loudest; compaction map kept -> original
//   k_stoi_tob      the silence-removed signals are never materialised: analysis frame g of the overlap-added
//                   kept frames is built from kept frames g-1, g, g+1 in shared memory, windowed again,
//                   transformed (512-point DFT of the 212 bins the 15 one-third-octave bands cover) and
//                   reduced to band magnitudes, for the clean and the processed signal
//   k_stoi_corr     30-frame segments: energy normalisation, clipping (-15 dB SDR bound), correlation
//   k_stoi_final    mean over segments and bands, pystoi's 1e-5 for fewer than 30 frames, optional running sums
#pragma once
#include <math.h>

#include "common.cuh"

namespace aw {

#define AW_STOI_FRAME 256
#define AW_STOI_HOP 128
#define AW_STOI_NFFT 512
#define AW_STOI_BANDS 15
#define AW_STOI_SEG 30
#define AW_STOI_BIN0 7           // first bin of band 0 (thirdoct(10 kHz, 512, 15, 150 Hz))
#define AW_STOI_NBIN 212         // bins 7 .. 218
#define AW_STOI_EPS 2.220446049250313e-16
#define AW_STOI_SEGS_PER_BLOCK 8

struct StoiArgs {
  const float* x; const float* y;     // clean / processed, 10 kHz
  long long sx, sy;
  int n;                  // samples per clip
  int F0;                 // frames of the silence gate: (n - 256) / 128 + 1
  double* energy;         // [clip][F0] dB
  unsigned long long* emax;  // [clip] order-preserving encoding of the largest energy
  int* src;               // [clip][F0] original frame of kept frame j
  int* kept;              // [clip] K
  float* tob;             // [clip][2][F0][16] band magnitudes of the analysis frames (clean, processed)
  double* part;           // [clip][seg blocks] partial correlation sums
  int seg_blocks;
  double* out;            // [clip]
  double* out_sum;        // [2] += {sum of scores > keep_above, their count} or null
  double keep_above;
};

__constant__ int c_stoi_edge[AW_STOI_BANDS + 1];   // band i = bins [edge[i], edge[i+1])

__device__ __forceinline__ float stoi_window(int m) {       // np.hanning(258)[1:-1]
  return 0.5f - 0.5f * cospif(2.0f * (float)(m + 1) / 257.0f);
}
__device__ __forceinline__ unsigned long long stoi_enc(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double stoi_dec(unsigned long long e) {
  const unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  return __longlong_as_double((long long)b);
}

// one warp per frame of the clean signal
__global__ void __launch_bounds__(256) k_stoi_energy(StoiArgs a) {
  const int clip = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * 8 + warp;
  if (f >= a.F0) return;
  const float* x = a.x + (long long)clip * a.sx + (long long)f * AW_STOI_HOP;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int m = lane + 32 * k;
    const double v = (double)(stoi_window(m) * x[m]);
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const double e = 20.0 * log10(sqrt(s) + AW_STOI_EPS);
    a.energy[(long long)clip * a.F0 + f] = e;
    atomicMax(a.emax + clip, stoi_enc(e));
  }
}

// one block per clip: mask = (max - 40 - e) < 0, exclusive scan, src[j] = f
__global__ void __launch_bounds__(256) k_stoi_scan(StoiArgs a) {
  __shared__ int s_w[8];
  __shared__ int s_base;
  const int clip = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double mx = stoi_dec(a.emax[clip]);
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int f0 = 0; f0 < a.F0; f0 += 256) {
    const int f = f0 + threadIdx.x;
    const bool keep = f < a.F0 && (mx - 40.0 - a.energy[(long long)clip * a.F0 + f]) < 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    if (keep) a.src[(long long)clip * a.F0 + off + __popc(bal & ((1u << lane) - 1u))] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += s_w[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) a.kept[clip] = s_base;
}

// one block per analysis frame g < K - 1 of the silence-removed signals
__global__ void __launch_bounds__(256) k_stoi_tob(StoiArgs a) {
  __shared__ float2 s_tw[AW_STOI_NFFT];
  __shared__ float s_sig[2][AW_STOI_FRAME];
  __shared__ float s_pow[2][AW_STOI_NBIN];
  const int clip = blockIdx.y, g = blockIdx.x, m = threadIdx.x;
  const int K = a.kept[clip];
  if (g >= K - 1) return;
  for (int i = m; i < AW_STOI_NFFT; i += 256) {
    float sn, cs;
    sincospif(2.0f * (float)i / (float)AW_STOI_NFFT, &sn, &cs);
    s_tw[i] = make_float2(cs, -sn);
  }
  const int* src = a.src + (long long)clip * a.F0;
  {
    // sample m of frame g of the overlap-added kept frames (hop 128, length 256: two frames overlap)
    const int fa = m < AW_STOI_HOP ? (g >= 1 ? src[g - 1] : -1) : src[g];
    const int fb = m < AW_STOI_HOP ? src[g] : src[g + 1];
    const int ma = m < AW_STOI_HOP ? m + AW_STOI_HOP : m;          // position inside frame fa
    const int mb = m < AW_STOI_HOP ? m : m - AW_STOI_HOP;          // position inside frame fb
    const float wa = stoi_window(ma), wb = stoi_window(mb), w2 = stoi_window(m);
    const float* x = a.x + (long long)clip * a.sx;
    const float* y = a.y + (long long)clip * a.sy;
    float vx = wb * x[(long long)fb * AW_STOI_HOP + mb], vy = wb * y[(long long)fb * AW_STOI_HOP + mb];
    if (fa >= 0) {
      vx += wa * x[(long long)fa * AW_STOI_HOP + ma];
      vy += wa * y[(long long)fa * AW_STOI_HOP + ma];
    }
    s_sig[0][m] = w2 * vx;
    s_sig[1][m] = w2 * vy;
  }
  __syncthreads();
  if (m < AW_STOI_NBIN) {
    const int k = AW_STOI_BIN0 + m;
    float xr = 0.f, xi = 0.f, yr = 0.f, yi = 0.f;
    int idx = 0;
#pragma unroll 8
    for (int t = 0; t < AW_STOI_FRAME; ++t) {
      const float2 tw = s_tw[idx];
      const float sx = s_sig[0][t], sy = s_sig[1][t];
      xr = fmaf(sx, tw.x, xr); xi = fmaf(sx, tw.y, xi);
      yr = fmaf(sy, tw.x, yr); yi = fmaf(sy, tw.y, yi);
      idx = (idx + k) & (AW_STOI_NFFT - 1);
    }
    s_pow[0][m] = xr * xr + xi * xi;
    s_pow[1][m] = yr * yr + yi * yi;
  }
  __syncthreads();
  if (m < 2 * 16) {
    const int sig = m >> 4, b = m & 15;
    float v = 0.f;
    if (b < AW_STOI_BANDS) {
      float s = 0.f;
      for (int k = c_stoi_edge[b]; k < c_stoi_edge[b + 1]; ++k) s += s_pow[sig][k - AW_STOI_BIN0];
      v = sqrtf(s);
    }
    a.tob[(((long long)clip * 2 + sig) * a.F0 + g) * 16 + b] = v;
  }
}

// thread = (segment, band); AW_STOI_SEGS_PER_BLOCK segments per block of 128 threads
__global__ void __launch_bounds__(128) k_stoi_corr(StoiArgs a) {
  __shared__ double s_red[4];
  const int clip = blockIdx.y;
  const int K = a.kept[clip], G = K - 1, J = G - AW_STOI_SEG + 1;
  const int seg = blockIdx.x * AW_STOI_SEGS_PER_BLOCK + (threadIdx.x >> 4), b = threadIdx.x & 15;
  double corr = 0.0;
  if (seg < J && b < AW_STOI_BANDS) {
    const float* tx = a.tob + (((long long)clip * 2 + 0) * a.F0 + seg) * 16 + b;
    const float* ty = a.tob + (((long long)clip * 2 + 1) * a.F0 + seg) * 16 + b;
    double xv[AW_STOI_SEG], yv[AW_STOI_SEG];
    double nx = 0.0, ny = 0.0;
#pragma unroll
    for (int i = 0; i < AW_STOI_SEG; ++i) {
      xv[i] = (double)tx[i * 16];
      yv[i] = (double)ty[i * 16];
      nx += xv[i] * xv[i];
      ny += yv[i] * yv[i];
    }
    const double c = sqrt(nx) / (sqrt(ny) + AW_STOI_EPS);
    const double clipv = 1.0 + 5.623413251903491;          // 1 + 10^(15/20)
    double my = 0.0, mxx = 0.0;
#pragma unroll
    for (int i = 0; i < AW_STOI_SEG; ++i) {
      yv[i] = fmin(yv[i] * c, xv[i] * clipv);
      my += yv[i];
      mxx += xv[i];
    }
    my /= AW_STOI_SEG;
    mxx /= AW_STOI_SEG;
    double sy = 0.0, sx = 0.0, sxy = 0.0;
#pragma unroll
    for (int i = 0; i < AW_STOI_SEG; ++i) {
      const double dy = yv[i] - my, dx = xv[i] - mxx;
      sy += dy * dy; sx += dx * dx; sxy += dy * dx;
    }
    corr = sxy / ((sqrt(sy) + AW_STOI_EPS) * (sqrt(sx) + AW_STOI_EPS));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) corr += __shfl_xor_sync(0xffffffffu, corr, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = corr;
  __syncthreads();
  if (threadIdx.x == 0)
    a.part[(long long)clip * a.seg_blocks + blockIdx.x] = s_red[0] + s_red[1] + s_red[2] + s_red[3];
}

__global__ void __launch_bounds__(128) k_stoi_final(StoiArgs a, int n_clips) {
  const int clip = blockIdx.x * blockDim.x + threadIdx.x;
  if (clip >= n_clips) return;
  const int G = a.kept[clip] - 1, J = G - AW_STOI_SEG + 1;
  double d = 1e-5;                                            // pystoi: "not enough frames", warns and returns 1e-5
  if (J >= 1) {
    double s = 0.0;
    const int nb = (J + AW_STOI_SEGS_PER_BLOCK - 1) / AW_STOI_SEGS_PER_BLOCK;
    for (int i = 0; i < nb; ++i) s += a.part[(long long)clip * a.seg_blocks + i];
    d = s / ((double)J * AW_STOI_BANDS);
  }
  a.out[clip] = d;
  if (a.out_sum && d > a.keep_above) {
    atomicAdd(a.out_sum, d);
    atomicAdd(a.out_sum + 1, 1.0);
  }
}

}  // namespace aw
