// Host-side constant tables of a plan: Hann window, Slaney mel filterbank (dense
// and the sparse two-slope layout the fused kernel consumes), orthonormal DCT-II
// rows, FFT twiddles and the steady-state initial conditions of an SOS cascade.
//
// These restate, in double precision with librosa's float32 rounding points,
//   librosa.filters.get_window('hann', fftbins=True) + util.pad_center,
//   librosa.filters.mel(htk=False, norm='slaney'),
//   scipy.fftpack.dct(type=2, norm='ortho') as a matrix,
//   scipy.signal.sosfilt_zi,
// which the reference reaches through script/mfcc.py:387 and :400-402.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "mmf_internal.h"

namespace mmf {

void host_window(int win_length, int n_fft, std::vector<float>& w) {
  w.assign(n_fft, 0.0f);
  const int lpad = (n_fft - win_length) / 2;
  for (int n = 0; n < win_length; ++n)
    w[lpad + n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)win_length));
}

static const double kFsp = 200.0 / 3.0;
static const double kMinLogHz = 1000.0;
static const double kMinLogMel = kMinLogHz / kFsp;
static const double kLogStep = std::log(6.4) / 27.0;

static double hz_to_mel(double f) {
  if (f >= kMinLogHz) return kMinLogMel + std::log(f / kMinLogHz) / kLogStep;
  return f / kFsp;
}
static double mel_to_hz(double m) {
  if (m >= kMinLogMel) return kMinLogHz * std::exp(kLogStep * (m - kMinLogMel));
  return kFsp * m;
}

void host_mel_dense(double sr, int n_fft, int n_mels, double fmin, double fmax, std::vector<float>& mel,
                    std::vector<double>& mel_f) {
  const int F = n_fft / 2 + 1;
  mel_f.resize(n_mels + 2);
  const double m_lo = hz_to_mel(fmin), m_hi = hz_to_mel(fmax);
  const int n = n_mels + 2;
  // numpy.linspace: start + i*step, last point set exactly
  const double step = (m_hi - m_lo) / (double)(n - 1);
  for (int i = 0; i < n; ++i) mel_f[i] = mel_to_hz(i == n - 1 ? m_hi : m_lo + (double)i * step);
  mel.assign((size_t)n_mels * F, 0.0f);
  const double rfft_val = 1.0 / ((double)n_fft * (1.0 / sr));  // np.fft.rfftfreq(n, d=1/sr): k * (1/(n*d))
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    for (int k = 0; k < F; ++k) {
      const double fk = (double)k * rfft_val;
      const double lower = -(mel_f[i] - fk) / fd0;
      const double upper = (mel_f[i + 2] - fk) / fd1;
      const double tri = std::max(0.0, std::min(lower, upper));
      const float tri32 = (float)tri;                  // stored into a float32 array ...
      mel[(size_t)i * F + k] = (float)((double)tri32 * enorm);  // ... then scaled in place
    }
  }
}

// Sparse layout: bin k belongs to segment seg(k) = #{centres mel_f[j] <= f_k} - 1
// and can only be non-zero in filters seg-1 (falling slope) and seg (rising slope).
bool host_mel_sparse(const std::vector<float>& mel, const std::vector<double>& mel_f, double sr, int n_fft, int n_mels,
                     MelSparse& out) {
  const int F = n_fft / 2 + 1;
  std::vector<int> seg(F);
  const double rfft_val = 1.0 / ((double)n_fft * (1.0 / sr));
  for (int k = 0; k < F; ++k) {
    const double fk = (double)k * rfft_val;
    int s = (int)(std::upper_bound(mel_f.begin(), mel_f.end(), fk) - mel_f.begin()) - 1;  // -1 .. n_mels+1
    seg[k] = s;
  }
  out.w2.assign((size_t)2 * F, 0.0f);
  for (int k = 0; k < F; ++k) {
    const int s = seg[k];
    for (int m = 0; m < n_mels; ++m) {
      const float w = mel[(size_t)m * F + k];
      if (w == 0.0f) continue;
      if (m == s - 1) {
        out.w2[2 * k] = w;
      } else if (m == s) {
        out.w2[2 * k + 1] = w;
      } else {
        return false;  // not a two-slope filterbank
      }
    }
  }
  out.seg_start.assign(n_mels + 2, F);
  for (int j = 0; j < n_mels + 2; ++j) {
    int first = F;
    for (int k = 0; k < F; ++k)
      if (seg[k] >= j) {
        first = k;
        break;
      }
    out.seg_start[j] = first;
  }
  return true;
}

// Grouped form of the sparse bank for the packed mel walk (stft_core.cuh: mel_groups): segment j is
// covered by groups of four consecutive bins aligned to multiples of four; the falling / rising weights of a
// group's bins are stored densely, zero outside the segment.
void host_mel_groups(const MelSparse& sp, int F, int n_mels, MelGroups& out) {
  out.segtab.assign((size_t)2 * (n_mels + 2), 0);
  out.segstep.assign((size_t)2 * (n_mels + 3), 0);
  out.w.clear();
  int widx = 0;
  int pos = 0;  // group the walk stands on after the previous segment
  for (int j = 0; j <= n_mels + 1; ++j) {
    const int a = sp.seg_start[j];
    const int b = j <= n_mels ? sp.seg_start[j + 1] : a;
    const int g0 = a / 4, g1 = b > a ? (b + 3) / 4 : g0;
    out.segtab[2 * j] = g0;
    out.segtab[2 * j + 1] = widx;
    out.segstep[2 * j] = g1 - g0;               // groups of segment j
    out.segstep[2 * j + 1] = j ? g0 - pos : 0;  // groups to step before it (-1 or 0)
    pos = g1;
    for (int g = g0; g < g1; ++g) {
      float dn[4], up[4];
      for (int i = 0; i < 4; ++i) {
        const int k = 4 * g + i;
        const bool in = k >= a && k < b && k < F;
        dn[i] = in ? sp.w2[2 * k] : 0.0f;
        up[i] = in ? sp.w2[2 * k + 1] : 0.0f;
      }
      out.w.insert(out.w.end(), dn, dn + 4);
      out.w.insert(out.w.end(), up, up + 4);
      ++widx;
    }
  }
  out.n_groups = widx;
}

void host_dct(int n_mfcc, int n_mels, std::vector<float>& d) {
  d.resize((size_t)n_mfcc * n_mels);
  for (int k = 0; k < n_mfcc; ++k)
    for (int m = 0; m < n_mels; ++m) {
      double v = k == 0 ? 1.0 / std::sqrt((double)n_mels)
                        : std::sqrt(2.0 / n_mels) * std::cos(M_PI * k * (2.0 * m + 1.0) / (2.0 * n_mels));
      d[(size_t)k * n_mels + m] = (float)v;
    }
}

void host_twiddles(int n_fft, const StftGeometry& g, std::vector<float2>& tw1, std::vector<float2>& tw2) {
  (void)n_fft;
  tw1.resize(g.tw1);
  tw2.resize(g.tw2 > 0 ? g.tw2 : 1);
  for (int k2 = 0; k2 < 16; ++k2)
    for (int n1 = 0; n1 < g.tpf; ++n1) {
      const double a = -2.0 * M_PI * (double)(((long)n1 * k2) % g.m) / (double)g.m;
      tw1[k2 * g.tpf + n1] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
  for (int j2 = 0; j2 < 16; ++j2)
    for (int m1 = 0; m1 < g.r3; ++m1) {
      const double a = -2.0 * M_PI * (double)((m1 * j2) % g.tpf) / (double)g.tpf;
      tw2[j2 * g.r3 + m1] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
}

// scipy.signal.sosfilt_zi + the default padlen of sosfiltfilt
int host_sos_zi(const double* sos, int n_sections, double* zi, int* padlen) {
  if (n_sections < 1 || n_sections > 16) return -1;
  double scale = 1.0;
  int zb = 0, za = 0;
  for (int s = 0; s < n_sections; ++s) {
    const double* c = sos + 6 * s;
    const double a0 = c[3];
    if (a0 == 0.0) return -1;
    const double b0 = c[0] / a0, b1 = c[1] / a0, b2 = c[2] / a0, a1 = c[4] / a0, a2 = c[5] / a0;
    // lfilter_zi: (I - A^T) zi = B with A = companion(a)^T, B = b[1:] - a[1:]*b[0]
    const double m00 = 1.0 + a1, m01 = -1.0, m10 = a2, m11 = 1.0;
    const double r0 = b1 - a1 * b0, r1 = b2 - a2 * b0;
    const double det = m00 * m11 - m01 * m10;
    if (det == 0.0) return -1;
    zi[2 * s] = scale * (r0 * m11 - m01 * r1) / det;
    zi[2 * s + 1] = scale * (m00 * r1 - m10 * r0) / det;
    scale *= (b0 + b1 + b2) / (1.0 + a1 + a2);
    if (c[2] == 0.0) ++zb;
    if (c[5] == 0.0) ++za;
  }
  const int ntaps = 2 * n_sections + 1 - std::min(zb, za);
  if (padlen) *padlen = 3 * ntaps;
  return 0;
}

}  // namespace mmf
