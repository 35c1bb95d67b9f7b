// Host emulator of the fused STFT kernel's per-thread phases (TEST INFRASTRUCTURE).
//
// Runs the exact __host__ __device__ phase functions of
// modulation_mfcc_b200/csrc/stft_core.cuh thread by thread on the CPU, with plain
// arrays standing in for shared memory and for warp shuffles.  It exists so the
// FFT index algebra can be checked against numpy.fft.rfft in the GPU-less
// container; it is not linked into the product library and nothing in the
// product path calls it.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../modulation_mfcc_b200/csrc/stft_core.cuh"

using namespace mmf;

// V = float2: one frame at a time; V = c2: frames (t, t+1) together through the packed-type path
template <int NFFT, typename V>
static void run(const float* y, long n, int hop, const float* window, float* power, long T, int regs_split) {
  using C = FftCfg<NFFT>;
  using TR = CxTraits<V>;
  using Tw = typename TR::Tw;
  using Xe = typename TR::Xe;
  constexpr int NF = TR::kFrames;
  const int pad = NFFT / 2;
  std::vector<float> ypad(n + 2 * pad + NFFT + (size_t)hop * NF, 0.0f);
  std::memcpy(ypad.data() + pad, y, sizeof(float) * n);
  // tables exactly as the plan builds them
  std::vector<Tw> tw1(C::TW1), tw2(C::TW2 > 0 ? C::TW2 : 1);
  for (int k2 = 0; k2 < 16; ++k2)
    for (int n1 = 0; n1 < C::TPF; ++n1) {
      double a = -2.0 * M_PI * (double)((long)n1 * k2 % C::M) / C::M;
      tw1[k2 * C::TPF + n1] = make_tw<V>(make_float2((float)cos(a), (float)sin(a)));
    }
  for (int j2 = 0; j2 < 16; ++j2)
    for (int m1 = 0; m1 < C::R3; ++m1) {
      double a = -2.0 * M_PI * (double)(m1 * j2 % C::TPF) / C::TPF;
      tw2[j2 * C::R3 + m1] = make_tw<V>(make_float2((float)cos(a), (float)sin(a)));
    }
  std::vector<Xe> xb(C::XBUF);
  const int pp = 2;  // power-tile pitch: two frames side by side
  std::vector<float> ptile((size_t)C::F * pp);
  std::vector<V> regs(C::TPF * 16);
  auto R = [&](int tau) -> V(&)[16] { return *reinterpret_cast<V(*)[16]>(&regs[tau * 16]); };
  auto wt = [&](int tau) {
    double a = -2.0 * M_PI * tau / NFFT;
    return make_tw<V>(make_float2((float)cos(a), (float)sin(a)));
  };
  // one exchange = write by all threads, then read by all threads; the two-frame path runs it per component
  auto xchg = [&](auto wr, auto rd) {
    for (int tau = 0; tau < C::TPF; ++tau) wr(tau);
    for (int tau = 0; tau < C::TPF; ++tau) rd(tau);
  };
  for (long t = 0; t < T; t += NF) {
    const float* span = ypad.data();
    const int off = (int)(t * hop);
    for (int tau = 0; tau < C::TPF; ++tau) {
      float2 wreg[16];
      for (int n2 = 0; n2 < 16; ++n2) {
        int c = tau + C::TPF * n2;
        wreg[n2] = make_float2(0.5f * window[2 * c], 0.5f * window[2 * c + 1]);
      }
      ph_load<NFFT, false>(R(tau), span, off, hop, tau, wreg);
      ph_pass1<NFFT>(R(tau), tw1.data(), tau);
    }
    if constexpr (NF == 1) {
      xchg([&](int tau) { ph_x1_write<NFFT, 0>(R(tau), xb.data(), tau); },
           [&](int tau) { ph_x1_read<NFFT, 0>(R(tau), xb.data(), tau); });
    } else {
      xchg([&](int tau) { ph_x1_write<NFFT, 1>(R(tau), xb.data(), tau); },
           [&](int tau) { ph_x1_read<NFFT, 1>(R(tau), xb.data(), tau); });
      xchg([&](int tau) { ph_x1_write<NFFT, 2>(R(tau), xb.data(), tau); },
           [&](int tau) { ph_x1_read<NFFT, 2>(R(tau), xb.data(), tau); });
    }
    for (int tau = 0; tau < C::TPF; ++tau) ph_pass2<NFFT>(R(tau), tw2.data(), tau);
    if (C::R3 > 1) {
      if constexpr (NF == 1) {
        xchg([&](int tau) { ph_x2_write<NFFT, 0>(R(tau), xb.data(), tau); },
             [&](int tau) { ph_x2_read<NFFT, 0>(R(tau), xb.data(), tau); });
      } else {
        xchg([&](int tau) { ph_x2_write<NFFT, 1>(R(tau), xb.data(), tau); },
             [&](int tau) { ph_x2_read<NFFT, 1>(R(tau), xb.data(), tau); });
        xchg([&](int tau) { ph_x2_write<NFFT, 2>(R(tau), xb.data(), tau); },
             [&](int tau) { ph_x2_read<NFFT, 2>(R(tau), xb.data(), tau); });
      }
      for (int tau = 0; tau < C::TPF; ++tau) ph_pass3<NFFT>(R(tau));
    }
    if (NFFT == 512 && regs_split) {
      for (int s = 0; s < 16; ++s) {
        V bpart[8];
        const int partner = (16 - s) & 15;
        for (int r = 0; r < 8; ++r) {
          V sh = R(partner)[15 - r];        // what the shuffles of round r deliver
          V own = R(s)[(16 - r) & 15];      // lane-0 special case
          bpart[r] = (s == 0) ? own : sh;
        }
        ph_split_regs512(R(s), bpart, ptile.data(), pp, 0, s, wt(s));
      }
    } else {
      std::vector<V> za((size_t)C::TPF * 9), zb((size_t)C::TPF * 8);
      auto ZA = [&](int tau) -> V(&)[9] { return *reinterpret_cast<V(*)[9]>(&za[(size_t)tau * 9]); };
      auto ZB = [&](int tau) -> V(&)[8] { return *reinterpret_cast<V(*)[8]>(&zb[(size_t)tau * 8]); };
      if constexpr (NF == 1) {
        xchg([&](int tau) { ph_z_write<NFFT, 0>(R(tau), xb.data(), tau); },
             [&](int tau) { ph_z_gather<NFFT, 0>(ZA(tau), ZB(tau), xb.data(), tau); });
      } else {
        xchg([&](int tau) { ph_z_write<NFFT, 1>(R(tau), xb.data(), tau); },
             [&](int tau) { ph_z_gather<NFFT, 1>(ZA(tau), ZB(tau), xb.data(), tau); });
        xchg([&](int tau) { ph_z_write<NFFT, 2>(R(tau), xb.data(), tau); },
             [&](int tau) { ph_z_gather<NFFT, 2>(ZA(tau), ZB(tau), xb.data(), tau); });
      }
      for (int tau = 0; tau < C::TPF; ++tau) ph_split_pairs<NFFT>(ZA(tau), ZB(tau), ptile.data(), pp, 0, tau, wt(tau));
    }
    for (int q = 0; q < NF && t + q < T; ++q)
      for (int k = 0; k < C::F; ++k) power[(long)k * T + t + q] = ptile[(size_t)k * pp + q];
  }
}

template <int NFFT>
static void run_any(const float* y, long n, int hop, const float* window, float* power, long T, int mode) {
  // mode bit 0: register split (n_fft 512); bit 1: two-frame packed-type path
  if (mode & 2)
    run<NFFT, c2>(y, n, hop, window, power, T, mode & 1);
  else
    run<NFFT, float2>(y, n, hop, window, power, T, mode & 1);
}

extern "C" int emu_stft_power(const float* y, long n, int nfft, int hop, const float* window, float* power, long T,
                              int mode) {
  switch (nfft) {
    case 32: run_any<32>(y, n, hop, window, power, T, mode); break;
    case 64: run_any<64>(y, n, hop, window, power, T, mode); break;
    case 128: run_any<128>(y, n, hop, window, power, T, mode); break;
    case 256: run_any<256>(y, n, hop, window, power, T, mode); break;
    case 512: run_any<512>(y, n, hop, window, power, T, mode); break;
    case 1024: run_any<1024>(y, n, hop, window, power, T, mode); break;
    case 2048: run_any<2048>(y, n, hop, window, power, T, mode); break;
    case 4096: run_any<4096>(y, n, hop, window, power, T, mode); break;
    default: return -1;
  }
  return 0;
}

// mel projection of a [F, T] power array with the sparse segment layout
extern "C" int emu_mel(const float* power, long T, int F, const int* seg_start, const float* w2, int n_mels,
                       int bands_per_worker, float* mel_out) {
  std::vector<float> col(F);
  for (long t = 0; t < T; ++t) {
    for (int k = 0; k < F; ++k) col[k] = power[(long)k * T + t];
    for (int m0 = 0; m0 < n_mels; m0 += bands_per_worker) {
      int m1 = m0 + bands_per_worker < n_mels ? m0 + bands_per_worker : n_mels;
      if ((m0 / bands_per_worker) & 1)
        mel_column<1>(col.data(), 1, seg_start, reinterpret_cast<const float2*>(w2), m0, m1,
                      [&](int m, float v) { mel_out[(long)m * T + t] = v; });
      else
        mel_column<0>(col.data(), 1, seg_start, reinterpret_cast<const float2*>(w2), m0, m1,
                      [&](int m, float v) { mel_out[(long)m * T + t] = v; });
    }
  }
  return 0;
}

// grouped mel walk (mel_groups) over the bin-pair tile layout (TilePairs): the power columns are stored
// exactly as the split step stores them (two-frame groups: put(k, t, pk) -> columns t/2 and t/2 + half; one-frame
// groups: put(k, t, float)), then every column is projected and un-permuted like the kernel does
extern "C" int emu_mel_groups(const float* power, long T, int F, const int* segtab, const int* segstep, const float* w,
                              int n_mels, int bands_per_worker, int tf, int two_frames, float* mel_out) {
  const int pp = tf + 2, half = tf / 2;
  const int rows = 2 * ((F + 3) / 4);  // bin-pair rows
  std::vector<float> tile((size_t)rows * pp * 2, 0.0f);
  for (long t0 = 0; t0 < T; t0 += tf) {
    std::fill(tile.begin(), tile.end(), 0.0f);
    TilePairs tp{tile.data(), pp, half};
    for (int t = 0; t < tf; t += two_frames ? 2 : 1)
      for (int k = 0; k < F; ++k) {
        auto P = [&](long tt) { return tt < T ? power[(long)k * T + tt] : 0.0f; };
        if (two_frames)
          tp.put(k, t, pmake(P(t0 + t), P(t0 + t + 1)));
        else
          tp.put(k, t, P(t0 + t));
      }
    for (int c = 0; c < tf; ++c) {
      const int t = two_frames ? ((c >= half ? c - half : c) << 1) + (c >= half ? 1 : 0) : c;
      if (t0 + t >= T) continue;
      for (int k = 0; k < F; ++k)
        if (tp.get(k, t, two_frames != 0) != power[(long)k * T + t0 + t]) return -2;
      for (int m0 = 0; m0 < n_mels; m0 += bands_per_worker) {
        const int m1 = m0 + bands_per_worker < n_mels ? m0 + bands_per_worker : n_mels;
        mel_groups(reinterpret_cast<const pk*>(tile.data()) + c, pp, reinterpret_cast<const int2*>(segtab),
                   reinterpret_cast<const int2*>(segstep), reinterpret_cast<const pk2*>(w), m0, m1,
                   [&](int m, float v) { mel_out[(long)m * T + t0 + t] = v; });
      }
    }
  }
  return 0;
}

// raw 16/8-point DFT checks
extern "C" void emu_dft16(const float* in, float* out) {
  float2 v[16];
  for (int i = 0; i < 16; ++i) v[i] = make_float2(in[2 * i], in[2 * i + 1]);
  dft16(v);
  for (int i = 0; i < 16; ++i) { out[2 * i] = v[i].x; out[2 * i + 1] = v[i].y; }
}
extern "C" void emu_dft8(const float* in, float* out) {
  float2 v[8];
  for (int i = 0; i < 8; ++i) v[i] = make_float2(in[2 * i], in[2 * i + 1]);
  dft8(v);
  for (int i = 0; i < 8; ++i) { out[2 * i] = v[i].x; out[2 * i + 1] = v[i].y; }
}
