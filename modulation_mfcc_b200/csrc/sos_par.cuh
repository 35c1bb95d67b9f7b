// Chunk-parallel zero-phase biquad cascade (scipy.signal.sosfiltfilt, used by the
// reference at script/mfcc.py:402, :421, :111) for rows that fit in shared memory.
//
// An IIR recurrence is sequential in time, but it is linear: a row of L samples
// is cut into 32 chunks of CL samples, one per lane.  Each lane
//   1. runs the cascade over its chunk from a zero state (lane 0: from the real
//      initial state zi*u[0]) and keeps only the final state,
//   2. the warp combines the 32 final states with a Kogge-Stone scan over the
//      affine maps  s -> M^(2^j) s + c  (M = zero-input transition of one chunk,
//      computed on the host in float64), which yields the exact state at every
//      chunk boundary,
//   3. each lane re-runs its chunk from its true initial state and writes the
//      output in place.
// The dependent chain is 2*CL + 5 steps instead of L, and every step is the same
// arithmetic as the sequential filter, so the result differs from scipy's only by
// float64 rounding of the carried states.
#pragma once
#include <cuda_runtime.h>

namespace mmf {

constexpr int kParMaxSections = 4;

struct SosPar {
  int ns;      // biquad sections (<= kParMaxSections)
  int padlen;  // scipy's default padlen for this cascade
  int CL;      // chunk length per lane (odd: conflict-free 64-bit shared accesses)
  int pad_;
  double sos[kParMaxSections][6];  // a0-normalised rows b0 b1 b2 1 a1 a2
  double zi[kParMaxSections][2];   // sosfilt_zi
  double mpow[5][2 * kParMaxSections][2 * kParMaxSections];  // zero-input transition over CL * 2^j samples
  double msb[2 * kParMaxSections][2 * kParMaxSections];      // ... over a whole super-block of 32 * CL samples
};

// cascade coefficients held in registers
template <int NS>
struct SosRegs {
  double b0[NS], b1[NS], b2[NS], a1[NS], a2[NS];
  __device__ __forceinline__ explicit SosRegs(const SosPar& a) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      b0[s] = a.sos[s][0];
      b1[s] = a.sos[s][1];
      b2[s] = a.sos[s][2];
      a1[s] = a.sos[s][4];
      a2[s] = a.sos[s][5];
    }
  }
};

// one sample through the cascade (direct form II transposed, scipy's _sosfilt order)
template <int NS>
__device__ __forceinline__ double sos_step(const SosRegs<NS>& c, double v, double (&z)[2 * NS]) {
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const double out = fma(c.b0[s], v, z[2 * s]);
    z[2 * s] = fma(c.b1[s], v, fma(-c.a1[s], out, z[2 * s + 1]));
    z[2 * s + 1] = fma(c.b2[s], v, -c.a2[s] * out);
    v = out;
  }
  return v;
}

// One direction over u[j] = REVERSE ? buf[L-1-j] : buf[j], j = 0..L-1, from initial state s0
// (meaningful on lane 0 only).  WRITE: outputs replace the inputs in place; otherwise only the
// state after the last full chunk grid (32 * CL samples, zero padded) is produced.  On return
// `z` holds the state at the end of this lane's chunk.  buf holds >= 32*CL doubles; all 32
// lanes of the warp must call.
template <int NS, bool REVERSE, bool WRITE>
__device__ __forceinline__ void warp_sos_core(double* buf, int L, const SosPar& a, const SosRegs<NS>& c, int lane,
                                              const double (&s0)[2 * NS], double (&z)[2 * NS]) {
  constexpr int D = 2 * NS;
  const int CL = a.CL;
  const int j0 = lane * CL;
  const int n = max(0, min(CL, L - j0));  // real samples of this chunk (the rest is zero padding)
  // element j of the processing order lives at base[j * step]
  double* base = REVERSE ? buf + (L - 1 - j0) : buf + j0;
  constexpr int step = REVERSE ? -1 : 1;
#pragma unroll
  for (int i = 0; i < D; ++i) z[i] = lane == 0 ? s0[i] : 0.0;
  // 1. final state of this chunk from a zero (lane 0: true) initial state
#pragma unroll 4
  for (int k = 0; k < n; ++k) sos_step<NS>(c, base[k * step], z);
  if (!WRITE)  // the carried state must cover the whole chunk grid: run the zero padding too
    for (int k = n; k < CL; ++k) sos_step<NS>(c, 0.0, z);
  // (WRITE: chunks past the end of the row carry garbage through the scan; nothing real depends on them)
  // 2. inclusive scan: z becomes the true state at the end of chunk `lane`
#pragma unroll
  for (int jj = 0; jj < 5; ++jj) {
    const int d = 1 << jj;
    double q[D];
#pragma unroll
    for (int i = 0; i < D; ++i) q[i] = __shfl_up_sync(0xffffffffu, z[i], d);
    if (lane >= d) {
      double zn[D];
#pragma unroll
      for (int r = 0; r < D; ++r) {
        double acc = z[r];
#pragma unroll
        for (int i = 0; i < D; ++i) acc = fma(a.mpow[jj][r][i], q[i], acc);
        zn[r] = acc;
      }
#pragma unroll
      for (int r = 0; r < D; ++r) z[r] = zn[r];
    }
  }
  if (WRITE) {
    // 3. true initial state of this chunk = end state of the previous one
    double s[D];
#pragma unroll
    for (int i = 0; i < D; ++i) s[i] = __shfl_up_sync(0xffffffffu, z[i], 1);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) s[i] = s0[i];
    }
#pragma unroll 4
    for (int k = 0; k < n; ++k) base[k * step] = sos_step<NS>(c, base[k * step], s);
  }
  __syncwarp();
}

// scipy's start-up: state = zi * first sample of the processing order
template <int NS, bool REVERSE>
__device__ __forceinline__ void warp_sos_pass(double* buf, int L, const SosPar& a, const SosRegs<NS>& c, int lane) {
  constexpr int D = 2 * NS;
  const double u0 = buf[REVERSE ? L - 1 : 0];
  double s0[D], z[D];
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    s0[2 * q] = a.zi[q][0] * u0;
    s0[2 * q + 1] = a.zi[q][1] * u0;
  }
  warp_sos_core<NS, REVERSE, true>(buf, L, a, c, lane, s0, z);
}

// forward + backward over the odd-extended row already sitting in buf[0..L)
template <int NS>
__device__ __forceinline__ void warp_sosfiltfilt(double* buf, int L, const SosPar& a, const SosRegs<NS>& c,
                                                 int lane) {
  warp_sos_pass<NS, false>(buf, L, a, c, lane);
  warp_sos_pass<NS, true>(buf, L, a, c, lane);
}

// odd extension (scipy.signal._arraytools.odd_ext) of row x[0..T) into buf[0..T+2p),
// computed in the input dtype as scipy does, then widened; buf[L..S) is zeroed.
template <typename TIn>
__device__ __forceinline__ void warp_odd_ext_staged(double* buf, int T, int p, int S, int lane) {
  // both extensions from the staged interior buf[p .. p+T) (values are exactly representable in TIn);
  // scipy's odd_ext does this arithmetic in the input dtype
  const int L = T + 2 * p;
  const TIn two = (TIn)2;
  const TIn x0 = (TIn)buf[p], xl = (TIn)buf[p + T - 1];
  for (int i = lane; i < p; i += 32) {
    buf[i] = (double)(TIn)(two * x0 - (TIn)buf[p + (p - i)]);
    buf[p + T + i] = (double)(TIn)(two * xl - (TIn)buf[p + T - 2 - i]);
  }
  for (int i = L + lane; i < S; i += 32) buf[i] = 0.0;
  __syncwarp();
}

template <typename TIn>
__device__ __forceinline__ void warp_load_odd_ext(const TIn* __restrict__ x, int T, int p, double* buf, int S,
                                                  int lane) {
  // interior: plain coalesced stream, eight loads in flight per lane.  The loads are unpredicated (index
  // clamped to the row) -- with a predicated remainder loop ptxas put every load next to its store and the
  // last 7 iterations of a 1001-sample row paid one memory latency each
  for (int t0 = 0; t0 < T; t0 += 256) {
    TIn a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = x[min(t0 + 32 * u + lane, T - 1)];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = t0 + 32 * u + lane;
      if (t < T) buf[p + t] = (double)a[u];
    }
  }
  __syncwarp();
  warp_odd_ext_staged<TIn>(buf, T, p, S, lane);
}

}  // namespace mmf
