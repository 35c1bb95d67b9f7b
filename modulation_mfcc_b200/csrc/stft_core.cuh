// Per-thread phases of the fused frame -> window -> real FFT -> |X|^2 path.
//
// Replaces the librosa stft under script/mfcc.py:387 (reference) for one frame.
// A frame of NFFT real samples is packed as M = NFFT/2 complex points
// z[c] = x[2c] + i*x[2c+1]; TPF = M/16 threads own 16 points each and run a
// mixed-radix Cooley-Tukey transform 16 x R2 x R3 with the data in registers and
// one shared-memory exchange between passes; a final split step turns Z into
// the real-input spectrum X[0..M] and stores |X|^2.
//
// Index algebra (all indices per frame; tau = thread within the frame):
//   n  = n1 + TPF*n2            n1 = tau,  n2 = 0..15      (register index)
//   pass 1: 16-pt DFT over n2 -> k2;  twiddle W_M^{n1*k2}
//   n1 = m1 + R3*m2             m1 < R3,   m2 < R2
//   pass 2: R2-pt DFT over m2 -> j2; twiddle W_TPF^{m1*j2}
//   pass 3: R3-pt DFT over m1 -> j1
//   k  = 16*(R2*j1 + j2) + k2
//
// Every phase is generic over the complex value type (fft_regs.cuh): float2 = one
// frame per thread, c2 = two adjacent frames per thread on packed FP32
// instructions.  The exchange buffer holds 8 bytes per element either way (the
// two-frame path exchanges real and imaginary parts in two rounds).
//
// Every phase is __host__ __device__: tests/emu/ runs the same code thread by
// thread on the CPU against numpy's rfft, so the index algebra is checked
// without a GPU.
#pragma once
#include "fft_regs.cuh"

namespace mmf {

template <int NFFT>
struct FftCfg {
  static_assert(NFFT >= 32 && NFFT <= 4096 && (NFFT & (NFFT - 1)) == 0, "n_fft must be a power of two in [32, 4096]");
  static constexpr int N = NFFT;
  static constexpr int M = NFFT / 2;          // complex points
  static constexpr int F = M + 1;             // real-spectrum bins
  static constexpr int TPF = M / 16;          // threads per frame
  static constexpr int R2 = TPF < 16 ? TPF : 16;
  static constexpr int R3 = TPF / R2;         // 1, 2, 4, 8
  static constexpr int NB2 = 16 / R2;         // pass-2 butterflies per thread
  static constexpr int NB3 = 16 / R3;         // pass-3 butterflies per thread
  static constexpr int PITCH1 = TPF + R3;     // exchange-1 row pitch (float2), conflict-free reads
  static constexpr int PITCH2 = 16 + 16 / R3; // exchange-2 row pitch (float2)
  static constexpr int X1 = 16 * PITCH1;
  static constexpr int X2 = R3 > 1 ? 16 * R3 * PITCH2 : 0;
  static constexpr int XZ = M;                // natural-order Z for the split step
  static constexpr int XBUF = (X1 > X2 ? (X1 > XZ ? X1 : XZ) : (X2 > XZ ? X2 : XZ));  // float2 per frame slot
  // distance between the exchange buffers of consecutive thread groups (8-byte elements): with
  // fewer than 16 threads per group several groups share a half-warp, and a stride that is a
  // multiple of 16 elements would put all of them on the same banks
  static constexpr int XSTRIDE = TPF >= 16 ? XBUF : XBUF + ((TPF % 16) - (XBUF % 16) + 16) % 16;
  static constexpr int TW1 = 16 * TPF;        // float2: W_M^{n1*k2} at [k2*TPF + n1]
  static constexpr int TW2 = 16 * R3;         // float2: W_TPF^{m1*j2} at [j2*R3 + m1]
};

// e^{-2*pi*i*r/32}, r = 0..8
MMF_HD float2 w32(int r) {
  constexpr float c[9] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                          0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f,
                          0.19509032201612825f, 0.0f};
  return make_float2(c[r], -c[8 - r]);
}

// component access for the exchanges.  PART 0: the whole value (float2 path, one
// round); PART 1 / 2: the real / imaginary pk of a two-frame value (two rounds
// through the same 8-byte-per-element buffer).
template <int PART>
MMF_HD float2 xget(const float2& v) {
  return v;
}
template <int PART>
MMF_HD void xset(float2& v, float2 e) {
  v = e;
}
template <int PART>
MMF_HD pk xget(const c2& v) {
  return PART == 1 ? v.x : v.y;
}
template <int PART>
MMF_HD void xset(c2& v, pk e) {
  if (PART == 1) {
    v.x = e;
  } else {
    v.y = e;
  }
}

// ---- phase L: gather 16 strided complex points of the frame and apply the window
// span: PCM of the tile in shared memory; frame starts at span[frame_off].
// wreg[n2] = 0.5 * (w[2c], w[2c+1]) with c = tau + TPF*n2 (the 0.5 is the 1/2 of
// the real-FFT split step, folded in exactly).
template <int NFFT, bool VEC>
MMF_HD void ph_load(float2 (&v)[16], const float* span, int frame_off, int hop, int tau, const float2 (&wreg)[16]) {
  using C = FftCfg<NFFT>;
  (void)hop;
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    const int c = tau + C::TPF * n2;
    float2 x;
    if constexpr (VEC) {
      x = *reinterpret_cast<const float2*>(span + frame_off + 2 * c);
    } else {
      x.x = span[frame_off + 2 * c];
      x.y = span[frame_off + 2 * c + 1];
    }
    v[n2] = make_float2(x.x * wreg[n2].x, x.y * wreg[n2].y);
  }
}

// Two frames (frame_off and frame_off + hop): windowed samples land directly in
// the packed halves (A.re, B.re), (A.im, B.im).
template <int NFFT, bool VEC>
MMF_HD void ph_load(c2 (&v)[16], const float* span, int frame_off, int hop, int tau, const float2 (&wreg)[16]) {
  using C = FftCfg<NFFT>;
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    const int c = tau + C::TPF * n2;
    float2 a, b;
    if constexpr (VEC) {
      a = *reinterpret_cast<const float2*>(span + frame_off + 2 * c);
      b = *reinterpret_cast<const float2*>(span + frame_off + hop + 2 * c);
    } else {
      a.x = span[frame_off + 2 * c];
      a.y = span[frame_off + 2 * c + 1];
      b.x = span[frame_off + hop + 2 * c];
      b.y = span[frame_off + hop + 2 * c + 1];
    }
    // scalar window products written straight into the packed halves (no regrouping moves)
    v[n2] = CxTraits<c2>::make(pmake(a.x * wreg[n2].x, b.x * wreg[n2].x), pmake(a.y * wreg[n2].y, b.y * wreg[n2].y));
  }
}

// Two frames whose hop is a multiple of 2*TPF samples (hop = 2*TPF*SH): point c of frame B is point c + TPF*SH of
// frame A, i.e. register n2 + SH of the SAME thread, so only SH of B's 16 points need loads of their own (21
// loads instead of 32 at hop 160, n_fft 512).  [lo, hi) is the range of n2 whose window values are not all zero
// (the window is zero-padded to n_fft: win 400 in 512 leaves n2 = 0 and 15 empty); points outside it are zero
// without touching the span.  Bitwise the same values as ph_load.
template <int NFFT, int SH, bool VEC>
MMF_HD void ph_load_shared(c2 (&v)[16], const float* span, int frame_off, int tau, const float2 (&wreg)[16], int lo,
                           int hi) {
  using C = FftCfg<NFFT>;
  float2 raw[16 + SH];
#pragma unroll
  for (int i = 0; i < 16 + SH; ++i) {
    const bool need = (i >= lo && i < hi) || (i - SH >= lo && i - SH < hi);
    raw[i] = make_float2(0.0f, 0.0f);
    if (need) {
      const int c = tau + C::TPF * i;
      if constexpr (VEC) {
        raw[i] = *reinterpret_cast<const float2*>(span + frame_off + 2 * c);
      } else {
        raw[i].x = span[frame_off + 2 * c];
        raw[i].y = span[frame_off + 2 * c + 1];
      }
    }
  }
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    const float2 a = raw[n2], b = raw[n2 + SH];
    v[n2] = CxTraits<c2>::make(pmake(a.x * wreg[n2].x, b.x * wreg[n2].x), pmake(a.y * wreg[n2].y, b.y * wreg[n2].y));
  }
}

// Same with first-order pre-emphasis y'[n] = y[n] - a*y[n-1] applied to the
// un-padded signal (y[-1] = 0) before the zero centre padding: n_valid is the
// number of samples from the frame start to the end of the clip, so samples at
// or beyond it stay exactly zero.  Needs one sample of history before the frame
// (the span is loaded with lead >= 1).
MMF_HD float2 pre_point(const float* span, int i0_abs, int i0, float a, long n_valid, float2 w) {
  const float xm = span[i0_abs - 1];
  const float x0 = span[i0_abs];
  const float x1 = span[i0_abs + 1];
  const float y0 = (i0 < n_valid) ? x0 - a * xm : 0.0f;
  const float y1 = (i0 + 1 < n_valid) ? x1 - a * x0 : 0.0f;
  return make_float2(y0 * w.x, y1 * w.y);
}
template <int NFFT>
MMF_HD void ph_load_pre(float2 (&v)[16], const float* span, int frame_off, int hop, int tau, const float2 (&wreg)[16],
                        float a, long n_valid) {
  using C = FftCfg<NFFT>;
  (void)hop;
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    const int i0 = 2 * (tau + C::TPF * n2);
    v[n2] = pre_point(span, frame_off + i0, i0, a, n_valid, wreg[n2]);
  }
}
template <int NFFT>
MMF_HD void ph_load_pre(c2 (&v)[16], const float* span, int frame_off, int hop, int tau, const float2 (&wreg)[16],
                        float a, long n_valid) {
  using C = FftCfg<NFFT>;
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    const int i0 = 2 * (tau + C::TPF * n2);
    const float2 fa = pre_point(span, frame_off + i0, i0, a, n_valid, wreg[n2]);
    const float2 fb = pre_point(span, frame_off + hop + i0, i0, a, n_valid - hop, wreg[n2]);
    v[n2] = CxTraits<c2>::make(pmake(fa.x, fb.x), pmake(fa.y, fb.y));
  }
}

// ---- pass 1: 16-point DFT over n2 and the W_M^{n1*k2} twiddle
template <int NFFT, typename V>
MMF_HD void ph_pass1(V (&v)[16], const typename CxTraits<V>::Tw* tw1, int tau) {
  using C = FftCfg<NFFT>;
  dft16(v);
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) v[k2] = cmul(v[k2], tw1[k2 * C::TPF + tau]);
}

// ---- exchange 1: A'[n1][k2] at xb[k2*PITCH1 + n1]
template <int NFFT, int PART = 0, typename V>
MMF_HD void ph_x1_write(const V (&v)[16], typename CxTraits<V>::Xe* xb, int tau) {
  using C = FftCfg<NFFT>;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) xb[k2 * C::PITCH1 + tau] = xget<PART>(v[k2]);
}

// thread tau' = m1 + R3*kq reads n1 = m1 + R3*m2, k2 = kq + R2*u into v[u*R2 + m2]
template <int NFFT, int PART = 0, typename V>
MMF_HD void ph_x1_read(V (&v)[16], const typename CxTraits<V>::Xe* xb, int tau) {
  using C = FftCfg<NFFT>;
  const int m1 = tau % C::R3, kq = tau / C::R3;
#pragma unroll
  for (int u = 0; u < C::NB2; ++u)
#pragma unroll
    for (int m2 = 0; m2 < C::R2; ++m2)
      xset<PART>(v[u * C::R2 + m2], xb[(kq + C::R2 * u) * C::PITCH1 + m1 + C::R3 * m2]);
}

// ---- pass 2: NB2 butterflies of radix R2 over m2 -> j2, twiddle W_TPF^{m1*j2}
template <int NFFT, typename V>
MMF_HD void ph_pass2(V (&v)[16], const typename CxTraits<V>::Tw* tw2, int tau) {
  using C = FftCfg<NFFT>;
  dft_groups<C::R2, C::NB2>(v);
  if constexpr (C::R3 > 1) {
    const int m1 = tau % C::R3;
#pragma unroll
    for (int j2 = 1; j2 < 16; ++j2) v[j2] = cmul(v[j2], tw2[j2 * C::R3 + m1]);
  }
}

// ---- exchange 2 (R3 > 1 only; R2 == 16, u == 0, kq == k2):
// B'[m1][j2][k2] at xb[(j2*R3 + m1)*PITCH2 + k2]
template <int NFFT, int PART = 0, typename V>
MMF_HD void ph_x2_write(const V (&v)[16], typename CxTraits<V>::Xe* xb, int tau) {
  using C = FftCfg<NFFT>;
  const int m1 = tau % C::R3, k2 = tau / C::R3;
#pragma unroll
  for (int j2 = 0; j2 < 16; ++j2) xb[(j2 * C::R3 + m1) * C::PITCH2 + k2] = xget<PART>(v[j2]);
}

// thread tau'' = k2 + 16*jq reads j2 = jq + R3*e, all m1, into v[e*R3 + m1]
template <int NFFT, int PART = 0, typename V>
MMF_HD void ph_x2_read(V (&v)[16], const typename CxTraits<V>::Xe* xb, int tau) {
  using C = FftCfg<NFFT>;
  const int k2 = tau % 16, jq = tau / 16;
#pragma unroll
  for (int e = 0; e < C::NB3; ++e)
#pragma unroll
    for (int m1 = 0; m1 < C::R3; ++m1)
      xset<PART>(v[e * C::R3 + m1], xb[((jq + C::R3 * e) * C::R3 + m1) * C::PITCH2 + k2]);
}

template <int NFFT, typename V>
MMF_HD void ph_pass3(V (&v)[16]) {
  using C = FftCfg<NFFT>;
  dft_groups<C::R3, C::NB3>(v);
}

// ---- natural-order store of Z for the split step: z[k], k = 16*(R2*j1 + j2) + k2
template <int NFFT, int PART = 0, typename V>
MMF_HD void ph_z_write(const V (&v)[16], typename CxTraits<V>::Xe* xb, int tau) {
  using C = FftCfg<NFFT>;
  if constexpr (C::R3 > 1) {
    const int k2 = tau % 16, jq = tau / 16;
#pragma unroll
    for (int e = 0; e < C::NB3; ++e)
#pragma unroll
      for (int j1 = 0; j1 < C::R3; ++j1) {
        const int j2 = jq + C::R3 * e;
        xb[16 * (C::R2 * j1 + j2) + k2] = xget<PART>(v[e * C::R3 + j1]);
      }
  } else {
    const int kq = tau;  // m1 == 0
#pragma unroll
    for (int u = 0; u < C::NB2; ++u)
#pragma unroll
      for (int j2 = 0; j2 < C::R2; ++j2) xb[16 * j2 + kq + C::R2 * u] = xget<PART>(v[u * C::R2 + j2]);
  }
}

// gather of the conjugate pairs this thread splits: a[r] = Z[k], b[r] = Z[(M-k) mod M]
// for k = tau + TPF*r, r = 0..7, and a[8] = Z[M/2] (used by tau == 0 only)
template <int NFFT, int PART = 0, typename V>
MMF_HD void ph_z_gather(V (&a)[9], V (&b)[8], const typename CxTraits<V>::Xe* xb, int tau) {
  using C = FftCfg<NFFT>;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int k = tau + C::TPF * r;
    xset<PART>(a[r], xb[k]);
    xset<PART>(b[r], xb[(C::M - k) & (C::M - 1)]);
  }
  xset<PART>(a[8], xb[C::M / 2]);
}

// ---- split step for one conjugate pair: a = Z[k], b = Z[(M-k) mod M] (both
// already scaled by 1/2 through the window).  Returns |X[k]|^2 in .x and
// |X[M-k]|^2 in .y (for two frames: each a pk of (frame A, frame B)).
// wk = e^{-2*pi*i*k/NFFT} is applied as c32[r] (compile-time) then wtau.
template <typename V>
MMF_HD V split_pair(V a, V b, float2 c32r, typename CxTraits<V>::Tw wtau) {
  const V E = cx<V>(sadd(a.x, b.x), ssub(a.y, b.y));
  const V O = cx<V>(sadd(a.y, b.y), ssub(b.x, a.x));
  const V WO = cmul(cmul(O, make_tw<V>(c32r)), wtau);
  const V Xp = cadd(E, WO), Xm = csub(E, WO);
  return cx<V>(sfma(Xp.x, Xp.x, smul(Xp.y, Xp.y)), sfma(Xm.x, Xm.x, smul(Xm.y, Xm.y)));
}

// power of bin k of frame t (float) / frames t, t+1 (pk; t even, ppitch even)
MMF_HD void ptile_store(float* ptile, int idx, float p) { ptile[idx] = p; }
MMF_HD void ptile_store(float* ptile, int idx, pk p) { *reinterpret_cast<pk*>(ptile + idx) = p; }

// ---- power-tile layouts (where the split step puts |X[k]|^2 of frame t)
// TileRows: [k][t], `pitch` floats per bin row -- read by the sparse mel walk (one frame per lane, one
// 32-bit load per bin) and by the mma.sync mel.
struct TileRows {
  float* base;
  int pitch;
  MMF_HD void put(int k, int t, float p) const { base[k * pitch + t] = p; }
  MMF_HD void put(int k, int t, pk p) const { *reinterpret_cast<pk*>(base + k * pitch + t) = p; }  // frames t, t+1
  MMF_HD float get(int k, int t) const { return base[k * pitch + t]; }
};
// TilePairs: [k / 2][column][k & 1] -- bins 2q and 2q+1 of ONE frame share an 8-byte word, so the mel
// phase reads two bins per 64-bit load and feeds them to one packed FFMA2 against two weights.  `pitch` =
// 8-byte words per bin-pair row.  Two-frame thread groups hold frames t (even) and t+1: they go to columns
// t/2 and t/2 + half (half = frames per tile / 2), which keeps the 32-bit stores of the two thread groups of
// a warp on disjoint banks when pitch = 2 (mod 16).
struct TilePairs {
  float* base;
  int pitch;
  int half;
  MMF_HD int col(int t, bool two_frames) const { return two_frames ? (t >> 1) + (t & 1) * half : t; }
  MMF_HD void put(int k, int t, float p) const { base[(((k >> 1) * pitch + t) << 1) + (k & 1)] = p; }
  MMF_HD void put(int k, int t, pk p) const {
    float* q = base + ((((k >> 1) * pitch) + (t >> 1)) << 1) + (k & 1);
    q[0] = plo(p);
    q[2 * half] = phi(p);
  }
  MMF_HD float get(int k, int t, bool two_frames) const { return base[(((k >> 1) * pitch + col(t, two_frames)) << 1) + (k & 1)]; }
};

// ---- split step from gathered pairs (any NFFT).
// Thread tau handles k = tau + TPF*r, r = 0..7 (covers [0, M/2)); tau == 0 also k = M/2.
// Power of bin k goes to st.put(k, t, power).
template <int NFFT, typename V, typename Tile>
MMF_HD void ph_split_pairs_to(const V (&a)[9], const V (&b)[8], const Tile& st, int t, int tau,
                              typename CxTraits<V>::Tw wtau) {
  using C = FftCfg<NFFT>;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int k = tau + C::TPF * r;
    const V p = split_pair(a[r], b[r], w32(r), wtau);
    st.put(k, t, p.x);
    st.put(C::M - k, t, p.y);
  }
  if (tau == 0) {
    const V p = split_pair(a[8], a[8], w32(8), wtau);
    st.put(C::M / 2, t, p.x);
  }
}
template <int NFFT, typename V>
MMF_HD void ph_split_pairs(const V (&a)[9], const V (&b)[8], float* ptile, int ppitch, int t, int tau,
                           typename CxTraits<V>::Tw wtau) {
  ph_split_pairs_to<NFFT>(a, b, TileRows{ptile, ppitch}, t, tau, wtau);
}

// one-frame convenience used by the host emulator and the trajectory-FFT kernel
template <int NFFT>
MMF_HD void ph_split_smem(const float2* xb, float* ptile, int ppitch, int t, int tau, float2 wtau) {
  float2 a[9], b[8];
  ph_z_gather<NFFT>(a, b, xb, tau);
  ph_split_pairs<NFFT>(a, b, ptile, ppitch, t, tau, wtau);
}

// Same, handing each (bin, power) to a callback instead of a power tile
// (used by the trajectory-FFT kernel of the modulation spectrum).
template <int NFFT, typename Emit>
MMF_HD void ph_split_smem_cb(const float2* xb, int tau, float2 wtau, Emit emit) {
  using C = FftCfg<NFFT>;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int k = tau + C::TPF * r;
    const float2 a = xb[k];
    const float2 b = xb[(C::M - k) & (C::M - 1)];
    const float2 p = split_pair(a, b, w32(r), wtau);
    emit(k, p.x);
    emit(C::M - k, p.y);
  }
  if (tau == 0) {
    const float2 a = xb[C::M / 2];
    const float2 p = split_pair(a, a, w32(8), wtau);
    emit(C::M / 2, p.x);
  }
}

// ---- split step for NFFT == 512 straight from registers.  After pass 2 thread
// s = tau holds Z[16*j2 + s] in v[j2].  bpart[r] must hold Z[M - (16*r + s)]:
// on the device it is v[15 - r] of lane (16 - s) & 15 (one shuffle per float);
// for s == 0 it is the thread's own v[(16 - r) & 15].
template <typename V, typename Tile>
MMF_HD void ph_split_regs512_to(const V (&v)[16], const V (&bpart)[8], const Tile& st, int t, int s,
                                typename CxTraits<V>::Tw wtau) {
  constexpr int M = 256;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int k = 16 * r + s;
    const V p = split_pair(v[r], bpart[r], w32(r), wtau);
    st.put(k, t, p.x);
    st.put(M - k, t, p.y);
  }
  if (s == 0) {
    const V p = split_pair(v[8], v[8], w32(8), wtau);
    st.put(M / 2, t, p.x);
  }
}
template <typename V>
MMF_HD void ph_split_regs512(const V (&v)[16], const V (&bpart)[8], float* ptile, int ppitch, int t, int s,
                             typename CxTraits<V>::Tw wtau) {
  ph_split_regs512_to(v, bpart, TileRows{ptile, ppitch}, t, s, wtau);
}

// ---- mel projection of one frame column from the power tile, for bands
// [m0, m1).  The Slaney filterbank (script/mfcc.py:387 -> librosa.filters.mel)
// is stored sparsely: bin k lies in segment seg(k) (between two filter centres)
// and feeds at most filter seg-1 (falling slope, w2[k].x) and filter seg
// (rising slope, w2[k].y).  seg_start[j] = first bin of segment j.
// pcol points at the column's bin 0 (bins are PITCH floats apart; PITCH = 0 takes
// the run-time ppitch).  Calls emit(m, value) for every band, in ascending order.
template <int PITCH, typename Emit>
MMF_HD void mel_column(const float* pcol, int ppitch, const int* seg_start, const float2* w2, int m0, int m1,
                       Emit emit) {
  const int pitch = PITCH > 0 ? PITCH : ppitch;
  float up_prev = 0.0f;
  int k = seg_start[m0];
  const float* pp = pcol + k * pitch;
  const float2* wp = w2 + k;
  int k_next = seg_start[m0 + 1];  // segment bounds are fetched one segment ahead of their use
#pragma unroll 1
  for (int j = m0; j <= m1; ++j) {
    const int n = k_next - k;
    k_next = seg_start[j + 2 <= m1 + 1 ? j + 2 : m1 + 1];
    k += n;
    float acc_dn = 0.0f, acc_up = 0.0f;
    int i = 0;
#pragma unroll 1
    for (; i + 2 <= n; i += 2) {
      const float p0 = pp[0], p1 = pp[pitch];
      const float2 wa = wp[0], wb = wp[1];
      acc_dn = fmaf(wa.x, p0, acc_dn);
      acc_up = fmaf(wa.y, p0, acc_up);
      acc_dn = fmaf(wb.x, p1, acc_dn);
      acc_up = fmaf(wb.y, p1, acc_up);
      pp += 2 * pitch;
      wp += 2;
    }
    if (i < n) {
      const float p0 = pp[0];
      const float2 wa = wp[0];
      acc_dn = fmaf(wa.x, p0, acc_dn);
      acc_up = fmaf(wa.y, p0, acc_up);
      pp += pitch;
      wp += 1;
    }
    if (j > m0) emit(j - 1, up_prev + acc_dn);
    up_prev = acc_up;
  }
}

// ---- mel projection, grouped form (TilePairs layout).  The two-slope bank is cut into segments as
// above; segment j is covered by n_j groups of four consecutive bins starting at bin 4*g0_j, with the
// falling / rising weights of the four bins stored densely (zero outside the segment), so that every
// irregularity of the filterbank sits in the weights and the loop body is uniform:
//   two 64-bit loads (bins 4g, 4g+1 | 4g+2, 4g+3 of this lane's frame), two broadcast 128-bit loads
//   (4 falling, 4 rising weights), four packed FFMA2 -- 8 multiply-adds in 8 instructions.
// segtab[j] = (g0_j, first weight group of segment j) gives a worker its starting point; from there the
// walk is incremental: segstep[j] = (n_j, pointer step into the tile before segment j, in pk units:
// -2*pp when segment j starts inside the last group of segment j-1, else 0) -- one 64-bit load per segment and
// no address is ever rebuilt.  w4[2*i] = falling weights of weight group i as two pk, w4[2*i+1] = rising weights.
// pcol = the lane's column of the tile (pk units), pp = pk per bin-pair row.
struct alignas(16) pk2 {
  pk a, b;
};
template <typename Emit>
MMF_HD void mel_groups(const pk* pcol, int pp, const int2* segtab, const int2* segstep, const pk2* w4, int m0, int m1,
                       Emit emit) {
  float up_prev = 0.0f;
  const int2 e0 = segtab[m0];
  const pk* p = pcol + (size_t)(2 * e0.x) * pp;
  const pk2* w = w4 + 2 * e0.y;
  const int2* st = segstep + m0;
  const int pp2 = 2 * pp;
  int n = st->x;  // the first segment starts at its own absolute position: no step
#pragma unroll 1
  for (int j = m0; j <= m1; ++j) {
    ++st;
    const int2 nx = *st;  // next segment's (count, step), fetched ahead of its use
    pk d0 = pmake(0.0f, 0.0f), d1 = d0, u0 = d0, u1 = d0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      const pk p01 = p[0], p23 = p[pp];
      const pk2 wd = w[0], wu = w[1];
      d0 = sfma(p01, wd.a, d0);
      d1 = sfma(p23, wd.b, d1);
      u0 = sfma(p01, wu.a, u0);
      u1 = sfma(p23, wu.b, u1);
      p += pp2;
      w += 2;
    }
    const pk d = sadd(d0, d1), u = sadd(u0, u1);
    if (j > m0) emit(j - 1, up_prev + (plo(d) + phi(d)));
    up_prev = plo(u) + phi(u);
    n = nx.x;
    p += nx.y;
  }
}

}  // namespace mmf
