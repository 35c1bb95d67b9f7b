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

template <int NFFT>
static void run(const float* y, long n, int hop, const float* window, float* power, long T, int regs_split) {
  using C = FftCfg<NFFT>;
  const int pad = NFFT / 2;
  std::vector<float> ypad(n + 2 * pad + NFFT, 0.0f);
  std::memcpy(ypad.data() + pad, y, sizeof(float) * n);
  // tables exactly as the plan builds them
  std::vector<float2> tw1(C::TW1), tw2(C::TW2 > 0 ? C::TW2 : 1);
  for (int k2 = 0; k2 < 16; ++k2)
    for (int n1 = 0; n1 < C::TPF; ++n1) {
      double a = -2.0 * M_PI * (double)((long)n1 * k2 % C::M) / C::M;
      tw1[k2 * C::TPF + n1] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int j2 = 0; j2 < 16; ++j2)
    for (int m1 = 0; m1 < C::R3; ++m1) {
      double a = -2.0 * M_PI * (double)(m1 * j2 % C::TPF) / C::TPF;
      tw2[j2 * C::R3 + m1] = make_float2((float)cos(a), (float)sin(a));
    }
  std::vector<float2> xb(C::XBUF);
  std::vector<float> ptile(C::F);
  std::vector<float2> regs(C::TPF * 16);
  auto R = [&](int tau) -> float2(&)[16] { return *reinterpret_cast<float2(*)[16]>(&regs[tau * 16]); };
  for (long t = 0; t < T; ++t) {
    const float* span = ypad.data();
    const int off = (int)(t * hop);
    for (int tau = 0; tau < C::TPF; ++tau) {
      float2 wreg[16];
      for (int n2 = 0; n2 < 16; ++n2) {
        int c = tau + C::TPF * n2;
        wreg[n2] = make_float2(0.5f * window[2 * c], 0.5f * window[2 * c + 1]);
      }
      ph_load<NFFT, false>(R(tau), span, off, tau, wreg);
      ph_pass1<NFFT>(R(tau), tw1.data(), tau);
      ph_x1_write<NFFT>(R(tau), xb.data(), tau);
    }
    for (int tau = 0; tau < C::TPF; ++tau) ph_x1_read<NFFT>(R(tau), xb.data(), tau);
    for (int tau = 0; tau < C::TPF; ++tau) ph_pass2<NFFT>(R(tau), tw2.data(), tau);
    if (C::R3 > 1) {
      for (int tau = 0; tau < C::TPF; ++tau) ph_x2_write<NFFT>(R(tau), xb.data(), tau);
      for (int tau = 0; tau < C::TPF; ++tau) ph_x2_read<NFFT>(R(tau), xb.data(), tau);
      for (int tau = 0; tau < C::TPF; ++tau) ph_pass3<NFFT>(R(tau));
    }
    if (NFFT == 512 && regs_split) {
      for (int s = 0; s < 16; ++s) {
        float2 bpart[8];
        const int partner = (16 - s) & 15;
        for (int r = 0; r < 8; ++r) {
          float2 sh = R(partner)[15 - r];        // what shuffle #r delivers
          float2 own = R(s)[(16 - r) & 15];      // lane-0 special case
          bpart[r] = (s == 0) ? own : sh;
        }
        double a = -2.0 * M_PI * s / NFFT;
        ph_split_regs512(R(s), bpart, ptile.data(), 1, 0, s, make_float2((float)cos(a), (float)sin(a)));
      }
    } else {
      for (int tau = 0; tau < C::TPF; ++tau) ph_z_write<NFFT>(R(tau), xb.data(), tau);
      for (int tau = 0; tau < C::TPF; ++tau) {
        double a = -2.0 * M_PI * tau / NFFT;
        ph_split_smem<NFFT>(xb.data(), ptile.data(), 1, 0, tau, make_float2((float)cos(a), (float)sin(a)));
      }
    }
    for (int k = 0; k < C::F; ++k) power[(long)k * T + t] = ptile[k];
  }
}

extern "C" int emu_stft_power(const float* y, long n, int nfft, int hop, const float* window, float* power, long T,
                              int regs_split) {
  switch (nfft) {
    case 32: run<32>(y, n, hop, window, power, T, regs_split); break;
    case 64: run<64>(y, n, hop, window, power, T, regs_split); break;
    case 128: run<128>(y, n, hop, window, power, T, regs_split); break;
    case 256: run<256>(y, n, hop, window, power, T, regs_split); break;
    case 512: run<512>(y, n, hop, window, power, T, regs_split); break;
    case 1024: run<1024>(y, n, hop, window, power, T, regs_split); break;
    case 2048: run<2048>(y, n, hop, window, power, T, regs_split); break;
    case 4096: run<4096>(y, n, hop, window, power, T, regs_split); break;
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
