// Attack simulations as streaming passes (reference
// scripts/attacks.py).  All randomness (start indices, band-stop edge, noise
// buffers) is drawn on the host and passed in, so runs are reproducible
// (the reference draws it unseeded: attacks.py:170,340,378).
#pragma once
#include "common.cuh"

namespace aw {

// ---- A1 PCMBitDepthConversion (attacks.py:44-70) ------------------------------
// a = x / max(|x| + 1e-8); q = trunc(clip(a * S, lo, hi)); out = float(q) / S
__global__ void __launch_bounds__(256) k_attack_pcm(const float* __restrict__ x, long long sx,
                                                    int n, const unsigned long long* peak,
                                                    float S, float lo, float hi,
                                                    float* __restrict__ out, long long so) {
  const int clip = blockIdx.y;
  const float d = peak_value(peak[clip]) + 1e-8f;
  const float* xc = x + (long long)clip * sx;
  float* oc = out + (long long)clip * so;
  int i0 = 0;
  if (((sx | so) & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int n4 = n >> 2;                                  // 16-byte streaming: 4 samples per access
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      const float4 t = reinterpret_cast<const float4*>(xc)[i];
      float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a = __fmul_rn(__fdiv_rn(v[k], d), S);
        a = fminf(fmaxf(a, lo), hi);
        v[k] = __fdiv_rn(truncf(a), S);
      }
      reinterpret_cast<float4*>(oc)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float a = __fmul_rn(__fdiv_rn(xc[i], d), S);
    a = fminf(fmaxf(a, lo), hi);
    oc[i] = __fdiv_rn(truncf(a), S);
  }
}

// ---- A2 Resample, sr // 16000 > 1 (attacks.py:276-287) ---------------------------
// decimate x[::f], then np.interp back (float64 arithmetic, last knot held)
__global__ void __launch_bounds__(256) k_attack_decim_interp(const float* __restrict__ x,
                                                             long long sx, int n, int f,
                                                             float* __restrict__ out, long long so) {
  const int clip = blockIdx.y;
  const float* p = x + (long long)clip * sx;
  float* oc = out + (long long)clip * so;
  const int last = ((n - 1) / f) * f;
  // np.interp: slope = (y1 - y0) / (x1 - x0).  For a power-of-two factor the division is exactly a
  // multiplication by 1/f (same rounding), which spares the ~30-instruction float64 division on a part
  // whose plain float64 rate is 1/64 of fp32.
  const bool pow2 = (f & (f - 1)) == 0;
  const double rf = 1.0 / (double)f;
  auto one = [&](int i) -> float {
    if (i >= last) return p[last];
    const int k0 = (i / f) * f;
    if (i == k0) return p[k0];                             // slope * 0 + y0 == y0 exactly
    const double y0 = p[k0], y1 = p[k0 + f];
    const double d = __dsub_rn(y1, y0);
    const double slope = pow2 ? __dmul_rn(d, rf) : __ddiv_rn(d, (double)f);
    return (float)__dadd_rn(__dmul_rn(slope, (double)(i - k0)), y0);
  };
  int i0 = 0;
  if ((so & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int n4 = n >> 2;                                  // 4 outputs per thread, one 16-byte store
    // One division per four outputs: the knot index k0 and the offset r advance incrementally, and a knot
    // value is loaded once (y1 of one interval is y0 of the next).  Same expression per output as `one`.
    const bool fast2 = f == 2 && (sx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      if (fast2 && 4 * i + 4 <= last) {                     // f = 2 (44.1 kHz): knots 4i, 4i+2, 4i+4 -- one 16-byte load + one
        const float4 q = reinterpret_cast<const float4*>(p)[i];   // scalar; slope * 1.0 is exact, so it is not computed
        const double a = q.x, b = q.z, c = p[4 * i + 4];
        const float v1 = (float)__dadd_rn(__dmul_rn(__dsub_rn(b, a), rf), a);
        const float v3 = (float)__dadd_rn(__dmul_rn(__dsub_rn(c, b), rf), b);
        reinterpret_cast<float4*>(oc)[i] = make_float4(q.x, v1, q.z, v3);
        continue;
      }
      int k0 = ((4 * i) / f) * f, r = 4 * i - k0;
      float y0 = p[min(k0, last)], y1 = (k0 + f <= last) ? p[k0 + f] : 0.f;
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k0 >= last) v[u] = p[last];
        else if (r == 0) v[u] = y0;
        else {
          const double d = __dsub_rn((double)y1, (double)y0);
          const double slope = pow2 ? __dmul_rn(d, rf) : __ddiv_rn(d, (double)f);
          v[u] = (float)__dadd_rn(__dmul_rn(slope, (double)r), (double)y0);
        }
        if (++r == f) { r = 0; k0 += f; y0 = y1; y1 = (k0 + f <= last) ? p[k0 + f] : 0.f; }
      }
      reinterpret_cast<float4*>(oc)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) oc[i] = one(i);
}

// ---- A2 Resample, polyphase branch (attacks.py:289-294): scipy upfirdn -----------
// y[k] = sum over x_i in [xi-hpp+1, xi] of x[x_i] * h_tf[phase*hpp + (x_i - (xi-hpp+1))],
// t = k*down, phase = t % up, xi = t / up, accumulated oldest sample first in float32
// (scipy/signal/_upfirdn_apply.pyx _apply_impl).  h_tf is the transposed+flipped,
// zero-padded filter prepared on the host; output k = k_off .. k_off + n_out - 1.
__global__ void __launch_bounds__(256) k_upfirdn(const float* __restrict__ x, long long sx, int n_in,
                                                 const float* __restrict__ h_tf, int hpp, int up,
                                                 int down, int k_off, int n_out,
                                                 float* __restrict__ out, long long so) {
  const int clip = blockIdx.y;
  const float* p = x + (long long)clip * sx;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_out; k += gridDim.x * blockDim.x) {
    const long long t = (long long)(k + k_off) * down;
    const int phase = (int)(t % up);
    const int xi = (int)(t / up);
    int x0 = xi - hpp + 1, hidx = phase * hpp;
    if (x0 < 0) { hidx -= x0; x0 = 0; }
    const int x1 = min(xi, n_in - 1);
    float acc = 0.f;
    for (int j = x0; j <= x1; ++j) acc = __fadd_rn(acc, __fmul_rn(p[j], h_tf[hidx++]));
    out[(long long)clip * so + k] = acc;
  }
}

// up = down = 1 (the FIR low / high / band-pass attacks): the same sums in the same order, register-tiled.
// Every thread owns FIR_R consecutive outputs and keeps a sliding window of FIR_R inputs in registers, so a
// tap costs one shared-memory load of x, one broadcast load of h and FIR_R multiply + add pairs (the products
// are NOT fused: scipy's float32 loop rounds the product before the add).  Samples outside [0, n_in) are staged
// as +0: acc + (+-0) == acc for every acc the loop can hold (acc is never -0), so that equals skipping them.
// The input tile is stored with one pad word per eight (index i + i/8): thread t reads word 9t + c.
#define AW_FIR_R 8
#define AW_FIR_TILE (256 * AW_FIR_R)
#define AW_FIR_MAXTAPS 1024
__device__ __forceinline__ int fir_addr(int i) { return i + (i >> 3); }

__global__ void __launch_bounds__(256) k_fir_tiled(const float* __restrict__ x, long long sx, int n_in,
                                                   const float* __restrict__ h_tf, int hpp, int k_off,
                                                   int n_out, float* __restrict__ out, long long so) {
  __shared__ float xs[((AW_FIR_TILE + AW_FIR_MAXTAPS + 8) * 9) / 8 + 8];
  __shared__ float hs[AW_FIR_MAXTAPS];
  const int clip = blockIdx.y;
  const float* p = x + (long long)clip * sx;
  float* o = out + (long long)clip * so;
  for (int j = threadIdx.x; j < hpp; j += 256) hs[j] = h_tf[j];
  for (int tile0 = blockIdx.x * AW_FIR_TILE; tile0 < n_out; tile0 += gridDim.x * AW_FIR_TILE) {
    const long long g0 = (long long)tile0 + k_off - hpp + 1;    // input index of staged word 0
    const int n_stage = AW_FIR_TILE + hpp + 7;
    __syncthreads();                                            // previous tile's output words are consumed
    for (int i = threadIdx.x; i < n_stage; i += 256) {
      const long long g = g0 + i;
      xs[fir_addr(i)] = (g >= 0 && g < n_in) ? p[g] : 0.f;
    }
    __syncthreads();
    const int base = threadIdx.x * AW_FIR_R;                    // output r sums xs[base + r + j] * hs[j], j ascending
    float acc[AW_FIR_R], w[AW_FIR_R];
#pragma unroll
    for (int r = 0; r < AW_FIR_R; ++r) { acc[r] = 0.f; w[r] = xs[fir_addr(base + r)]; }
    int j = 0;
    for (; j + 8 <= hpp; j += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float h = hs[j + u];
#pragma unroll
        for (int r = 0; r < AW_FIR_R; ++r) acc[r] = __fadd_rn(acc[r], __fmul_rn(w[(u + r) & 7], h));
        w[u] = xs[fir_addr(base + j + u + 8)];                  // x[base + j + u] is done; its slot takes x[.. + 8]
      }
    }
    for (; j < hpp; ++j) {
      const float h = hs[j];
#pragma unroll
      for (int r = 0; r < AW_FIR_R; ++r) acc[r] = __fadd_rn(acc[r], __fmul_rn(xs[fir_addr(base + j + r)], h));
    }
    __syncthreads();                                            // all reads of the input tile are done
#pragma unroll
    for (int r = 0; r < AW_FIR_R; ++r) xs[fir_addr(base + r)] = acc[r];
    __syncthreads();
    const int n_here = min(AW_FIR_TILE, n_out - tile0);
    for (int i = threadIdx.x; i < n_here; i += 256) o[tile0 + i] = xs[fir_addr(i)];
  }
}

// ---- A3/A4/A5 Butterworth IIR (attacks.py:342-349, 413-416, 451-453) ---------------
// scipy lfilter = direct form II transposed in float64.  The recurrence is made
// parallel by chunking: every thread owns one chunk and first runs `warm` samples
// of look-back from zero state; the filter's memory of anything older has decayed
// below double rounding (warm is chosen on the host from the largest pole radius),
// so the state at the chunk start is the sequential one to ~1e-16 relative.  The
// first chunk starts from the exact initial state.
#define AW_IIR_MAXORD 8
enum { IIR_SRC_F32 = 0, IIR_SRC_ODDEXT = 1, IIR_SRC_REV_F64 = 2 };
enum { IIR_DST_F32 = 0, IIR_DST_F64 = 1, IIR_DST_REVTRIM_F32 = 2 };

struct IirArgs {
  double b[AW_IIR_MAXORD + 1], a[AW_IIR_MAXORD + 1], zi[AW_IIR_MAXORD];
  int order;
  int n;          // number of samples filtered (incl. extension)
  int n_clips, n_chunks;
  int chunk, warm;
  int padlen;     // ODDEXT / REVTRIM: edge extension length (27)
  int use_zi;     // initial state = zi * first sample (filtfilt) else 0
  const float* x32; long long sx32; int n_x;   // original signal
  const double* x64; long long sx64;           // REV_F64 source (length n)
  float* o32; long long so32;
  double* o64; long long so64;
};

template <int SRC>
__device__ __forceinline__ double iir_load(const IirArgs& a, int clip, int i) {
  if (SRC == IIR_SRC_F32) return (double)a.x32[(long long)clip * a.sx32 + i];
  if (SRC == IIR_SRC_REV_F64) return a.x64[(long long)clip * a.sx64 + (a.n - 1 - i)];
  // odd extension: [2 x0 - x[pad..1], x, 2 x_last - x[n-2 .. n-pad-1]]
  const float* p = a.x32 + (long long)clip * a.sx32;
  const int j = i - a.padlen;
  if (j < 0) return __dsub_rn(__dmul_rn(2.0, (double)p[0]), (double)p[-j]);
  if (j >= a.n_x) return __dsub_rn(__dmul_rn(2.0, (double)p[a.n_x - 1]), (double)p[2 * (a.n_x - 1) - j]);
  return (double)p[j];
}

// One thread per (clip, chunk).  n_chunks == 1 is the sequential mode: one thread walks a whole
// clip from the exact initial state and reproduces scipy's recurrence bit for bit (needed for
// the direct-form band-stop, whose own round-off noise reaches 1e-6..1e-3: any re-start of the
// recurrence lands on a different noise realisation).
template <int SRC, int DST>
__global__ void __launch_bounds__(128) k_iir(IirArgs a) {
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long long)a.n_clips * a.n_chunks) return;
  const int clip = (int)(item / a.n_chunks);
  const int ch = (int)(item - (long long)clip * a.n_chunks);
  const int start = ch * a.chunk;
  if (start >= a.n) return;
  const int end = min(start + a.chunk, a.n);
  double z[AW_IIR_MAXORD];
#pragma unroll
  for (int k = 0; k < AW_IIR_MAXORD; ++k) z[k] = 0.0;
  int i = start - a.warm;
  if (i <= 0) {
    i = 0;
    if (a.use_zi) {
      const double x0 = iir_load<SRC>(a, clip, 0);
#pragma unroll
      for (int k = 0; k < AW_IIR_MAXORD; ++k) z[k] = k < a.order ? __dmul_rn(a.zi[k], x0) : 0.0;
    }
  }
  for (; i < end; ++i) {
    const double x = iir_load<SRC>(a, clip, i);
    const double y = __dadd_rn(z[0], __dmul_rn(a.b[0], x));
#pragma unroll
    for (int k = 0; k < AW_IIR_MAXORD; ++k) {
      if (k < a.order) {
        const double zn = k + 1 < a.order ? z[k + 1 < AW_IIR_MAXORD ? k + 1 : 0] : 0.0;
        z[k] = __dsub_rn(__dadd_rn(zn, __dmul_rn(x, a.b[k + 1])), __dmul_rn(y, a.a[k + 1]));
      }
    }
    if (i >= start) {
      if (DST == IIR_DST_F32) a.o32[(long long)clip * a.so32 + i] = (float)y;
      if (DST == IIR_DST_F64) a.o64[(long long)clip * a.so64 + i] = y;
      if (DST == IIR_DST_REVTRIM_F32) {
        const int j = (a.n - 1 - i) - a.padlen;      // un-reverse, drop the extension
        if (j >= 0 && j < a.n_x) a.o32[(long long)clip * a.so32 + j] = (float)y;
      }
    }
  }
}

// Sequential mode, one thread per clip, software-pipelined: the recurrence itself is a chain of three
// dependent float64 operations per sample (~30 cycles), but a naive loop also waits a full,
// uncoalesced memory latency per sample (measured 380 ns / sample).  Here the next 32 inputs are
// requested while the current 16 are filtered, and outputs leave in batches, so the loop runs at the
// speed of the dependency chain.  The arithmetic (operation order, no FMA contraction) is unchanged:
// bit-identical to scipy's lfilter.
template <int SRC, int DST>
__global__ void __launch_bounds__(32) k_iir_seq(IirArgs a) {
  const int clip = blockIdx.x * blockDim.x + threadIdx.x;
  if (clip >= a.n_clips) return;
  constexpr int NB = 16;                 // 3 x 16 doubles of staging: no register spills
  double z[AW_IIR_MAXORD];
#pragma unroll
  for (int k = 0; k < AW_IIR_MAXORD; ++k) z[k] = 0.0;
  if (a.use_zi) {
    const double x0 = iir_load<SRC>(a, clip, 0);
#pragma unroll
    for (int k = 0; k < AW_IIR_MAXORD; ++k) z[k] = k < a.order ? __dmul_rn(a.zi[k], x0) : 0.0;
  }
  double xc[NB], xn[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) xc[j] = j < a.n ? iir_load<SRC>(a, clip, j) : 0.0;
  for (int i0 = 0; i0 < a.n; i0 += NB) {
#pragma unroll
    for (int j = 0; j < NB; ++j) xn[j] = i0 + NB + j < a.n ? iir_load<SRC>(a, clip, i0 + NB + j) : 0.0;
    double yb[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const double x = xc[j];
      const double y = __dadd_rn(z[0], __dmul_rn(a.b[0], x));
#pragma unroll
      for (int k = 0; k < AW_IIR_MAXORD; ++k) {
        if (k < a.order) {
          const double zn = k + 1 < a.order ? z[k + 1 < AW_IIR_MAXORD ? k + 1 : 0] : 0.0;
          z[k] = __dsub_rn(__dadd_rn(zn, __dmul_rn(x, a.b[k + 1])), __dmul_rn(y, a.a[k + 1]));
        }
      }
      yb[j] = y;
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int i = i0 + j;
      if (i >= a.n) break;
      if (DST == IIR_DST_F32) a.o32[(long long)clip * a.so32 + i] = (float)yb[j];
      if (DST == IIR_DST_F64) a.o64[(long long)clip * a.so64 + i] = yb[j];
      if (DST == IIR_DST_REVTRIM_F32) {
        const int jj = (a.n - 1 - i) - a.padlen;
        if (jj >= 0 && jj < a.n_x) a.o32[(long long)clip * a.so32 + jj] = (float)yb[j];
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) xc[j] = xn[j];
  }
}

// Sequential mode, EIGHT LANES PER CLIP.  Plain (non-tensor) float64 issues at ~2 lanes per scheduler per
// clock on this part (measured: 16 cycles per warp instruction), so one thread per clip spends
// 34 x 16 cycles per sample on the eight state updates.  Here lane k of an 8-lane group owns z[k] (and
// b[k+1], a[k+1]): per sample the group computes y on lane 0, broadcasts it, and every lane updates its
// own state -- 6 float64 warp instructions per sample instead of 34, four clips per warp, 64 warps for
// 256 clips.  Every operation, operand and rounding is the one the single-thread recurrence (and scipy's
// lfilter) performs, so the result is bit-identical; orders below 8 run with zero coefficients in the
// unused lanes (their state stays exactly 0).
template <int SRC, int DST>
__global__ void __launch_bounds__(32) k_iir_seq8(IirArgs a) {
  const int lane = threadIdx.x, g = lane >> 3, k = lane & 7;
  const int clip_raw = blockIdx.x * 4 + g;
  const bool live = clip_raw < a.n_clips;
  const int clip = live ? clip_raw : a.n_clips - 1;
  const int base = g * 8;
  const double bk = k + 1 <= a.order ? a.b[k + 1] : 0.0, ak = k + 1 <= a.order ? a.a[k + 1] : 0.0;
  const double b0 = a.b[0];
  double z = 0.0;
  if (a.use_zi) z = k < a.order ? __dmul_rn(a.zi[k], iir_load<SRC>(a, clip, 0)) : 0.0;
  double xv = k < a.n ? iir_load<SRC>(a, clip, k) : 0.0;          // lane k holds sample i0 + k of the batch
  for (int i0 = 0; i0 < a.n; i0 += 8) {
    const double xn = i0 + 8 + k < a.n ? iir_load<SRC>(a, clip, i0 + 8 + k) : 0.0;    // next batch, in flight
    double ykeep = 0.0;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const double x = __shfl_sync(0xffffffffu, xv, base + s);
      const double y0 = __dadd_rn(z, __dmul_rn(b0, x));             // meaningful on lane 0 of the group
      const double y = __shfl_sync(0xffffffffu, y0, base);
      double zn = __shfl_down_sync(0xffffffffu, z, 1);              // z[k + 1] of the same group
      if (k == 7) zn = 0.0;
      if (i0 + s < a.n) z = __dsub_rn(__dadd_rn(zn, __dmul_rn(x, bk)), __dmul_rn(y, ak));
      if (k == s) ykeep = y;
    }
    const int i = i0 + k;
    if (live && i < a.n) {
      if (DST == IIR_DST_F32) a.o32[(long long)clip * a.so32 + i] = (float)ykeep;
      if (DST == IIR_DST_F64) a.o64[(long long)clip * a.so64 + i] = ykeep;
      if (DST == IIR_DST_REVTRIM_F32) {
        const int jj = (a.n - 1 - i) - a.padlen;
        if (jj >= 0 && jj < a.n_x) a.o32[(long long)clip * a.so32 + jj] = (float)ykeep;
      }
    }
    xv = xn;
  }
}

// ---- A6 DeleteSamples / A7 SampleSupression / A8 Cropout (attacks.py:162-205,370-385)
__global__ void __launch_bounds__(256) k_attack_delete(const float* __restrict__ x, long long sx,
                                                       int n_out, const int* __restrict__ start,
                                                       int n_del, float* __restrict__ out,
                                                       long long so) {
  const int clip = blockIdx.y, s = start[clip];
  const float* xc = x + (long long)clip * sx;
  float* oc = out + (long long)clip * so;
  int i0 = 0;
  if ((so & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int n4 = n_out >> 2;                              // 16-byte stores; the shifted loads stay scalar (coalesced)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      const int j = 4 * i;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = xc[j + k < s ? j + k : j + k + n_del];
      reinterpret_cast<float4*>(oc)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x)
    oc[i] = xc[i < s ? i : i + n_del];
}

__global__ void __launch_bounds__(256) k_attack_suppress(const float* __restrict__ x, long long sx,
                                                         int n, const int* __restrict__ start,
                                                         int n_zero, float* __restrict__ out,
                                                         long long so) {
  const int clip = blockIdx.y, s = start[clip];
  const float* xc = x + (long long)clip * sx;
  float* oc = out + (long long)clip * so;
  int i0 = 0;
  if (((sx | so) & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      const int j = 4 * i;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j + 3 < s || j >= s + n_zero) {
        t = reinterpret_cast<const float4*>(xc)[i];
      } else if (!(j >= s && j + 3 < s + n_zero)) {          // straddles an edge of the zeroed span
        t = reinterpret_cast<const float4*>(xc)[i];
        if (j >= s && j < s + n_zero) t.x = 0.f;
        if (j + 1 >= s && j + 1 < s + n_zero) t.y = 0.f;
        if (j + 2 >= s && j + 2 < s + n_zero) t.z = 0.f;
        if (j + 3 >= s && j + 3 < s + n_zero) t.w = 0.f;
      }
      reinterpret_cast<float4*>(oc)[i] = t;
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    oc[i] = (i >= s && i < s + n_zero) ? 0.f : xc[i];
}

// ---- extensions with no reference counterpart ("parity unpinned") ------
// y = gain * x + sigma * noise   (noise: host-seeded buffer or null)
__global__ void __launch_bounds__(256) k_attack_affine(const float* __restrict__ x, long long sx,
                                                       int n, float gain,
                                                       const float* __restrict__ noise,
                                                       long long sn, float sigma,
                                                       float* __restrict__ out, long long so) {
  const int clip = blockIdx.y;
  const float* xc = x + (long long)clip * sx;
  const float* nc = noise ? noise + (long long)clip * sn : nullptr;
  float* oc = out + (long long)clip * so;
  int i0 = 0;
  uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out);
  long long st = sx | so;
  if (noise) { al |= reinterpret_cast<uintptr_t>(noise); st |= sn; }
  if ((st & 3) == 0 && (al & 15) == 0) {                      // 16-byte streaming: 4 samples per access
    const int n4 = n >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(xc) + i);
      float4 v = make_float4(__fmul_rn(gain, t.x), __fmul_rn(gain, t.y), __fmul_rn(gain, t.z), __fmul_rn(gain, t.w));
      if (nc) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(nc) + i);
        v.x = __fadd_rn(v.x, __fmul_rn(sigma, q.x)); v.y = __fadd_rn(v.y, __fmul_rn(sigma, q.y));
        v.z = __fadd_rn(v.z, __fmul_rn(sigma, q.z)); v.w = __fadd_rn(v.w, __fmul_rn(sigma, q.w));
      }
      reinterpret_cast<float4*>(oc)[i] = v;
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float v = __fmul_rn(gain, xc[i]);
    if (nc) v = __fadd_rn(v, __fmul_rn(sigma, nc[i]));
    oc[i] = v;
  }
}

// ---- X3 "compression approximation" (named by the build brief; NO reference arithmetic: upstream's
// MP3 attack shells out to ffmpeg, attacks.py:73-148 -- parity unpinned, defined here) -------------
// A transform codec in one line: coarse log-magnitude quantisation plus masking of weak bins, applied
// to the embedding band of the STFT and resynthesised with the original phase.  Per frame:
//     floor = max_b |X_b| * 10^(floor_db/20);   |X_b| < floor -> 0;
//     else  |X_b| -> 10^(step_db * rint(20 log10|X_b| / step_db) / 20)
// The kernel emits the magnitude CHANGE (q - |X|): the caller adds iSTFT(change * phasor) to the input,
// so everything outside the band passes through untouched.  One warp per frame.
__global__ void __launch_bounds__(128) k_spectral_quantize(const float* __restrict__ mag, long long frames,
                                                           int nb, float k_log, float k_exp, float floor_ratio,
                                                           float* __restrict__ dmag) {
  const long long f = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= frames) return;
  const float* src = mag + f * nb;
  float v[8];                                            // nb <= AW_MAX_BINS = 256
  float mx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int b = lane + 32 * i;
    v[i] = b < nb ? src[b] : 0.f;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  const float floor_v = __fmul_rn(mx, floor_ratio);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int b = lane + 32 * i;
    if (b >= nb) continue;
    float q = 0.f;
    if (v[i] >= floor_v && v[i] > 0.f) q = exp2f(__fmul_rn(rintf(__fmul_rn(log2f(v[i]), k_log)), k_exp));
    dmag[f * nb + b] = __fsub_rn(q, v[i]);
  }
}

}  // namespace aw
