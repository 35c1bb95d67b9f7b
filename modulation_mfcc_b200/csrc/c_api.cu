// extern "C" surface of libmmf_b200.so (declared in include/mmf.h).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cudaTypedefs.h>

#include "mmf_internal.h"
#include "sos_par.cuh"

namespace mmf {

static thread_local std::string g_err;
static thread_local long g_launches = 0;

void set_error(const std::string& msg) { g_err = msg; }
void count_launch(int n) { g_launches += n; }

static int fail(int code, const std::string& msg) {
  set_error(msg);
  return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
  return fail(MMF_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define MMF_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

static int validate_cfg(const mmf_config* c) {
  if (!c) return fail(MMF_ERR_INVALID, "config is NULL");
  if (c->n_fft < 16 || c->n_fft > 4096) return fail(MMF_ERR_UNSUPPORTED, "n_fft must be in [16, 4096]");
  if (c->win_length < 1 || c->win_length > c->n_fft)
    return fail(MMF_ERR_INVALID, "Target size (n_fft) must be at least input size (win_length)");
  if (c->hop_length < 1) return fail(MMF_ERR_INVALID, "hop_length must be a positive integer");
  if (c->n_mels < 1 || c->n_mels > 512) return fail(MMF_ERR_UNSUPPORTED, "n_mels must be in [1, 512]");
  if (c->n_mfcc < 1 || c->n_mfcc > 128 || c->n_mfcc > c->n_mels)
    return fail(MMF_ERR_UNSUPPORTED, "n_mfcc must be in [1, min(128, n_mels)]");
  if (!(c->sample_rate > 0)) return fail(MMF_ERR_INVALID, "sample_rate must be positive");
  if (!(c->fmax > c->fmin) || c->fmin < 0) return fail(MMF_ERR_INVALID, "need 0 <= fmin < fmax");
  return MMF_OK;
}

template <typename T>
static cudaError_t upload(T** dst, const std::vector<T>& src) {
  cudaError_t e = cudaMalloc((void**)dst, std::max<size_t>(1, src.size()) * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static int ensure_ws(mmf_plan* p, size_t bytes) {
  if (bytes <= p->ws_bytes) return MMF_OK;
  if (p->ws) {
    MMF_CUDA(cudaDeviceSynchronize());
    MMF_CUDA(cudaFree(p->ws));
    p->ws = nullptr;
    p->ws_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&p->ws, bytes);
  if (e != cudaSuccess) return fail(MMF_ERR_NOMEM, std::string("workspace cudaMalloc: ") + cudaGetErrorString(e));
  p->ws_bytes = bytes;
  return MMF_OK;
}

// The host-buffer entry points run on the plan's own streams: they get a workspace of their own so that
// they never alias buffers of device-resident calls still in flight on the caller's stream.
static int ensure_host_ws(mmf_plan* p, size_t bytes) {
  if (bytes <= p->host_ws_bytes) return MMF_OK;
  if (p->host_ws) {
    MMF_CUDA(cudaStreamSynchronize(p->streams[0]));
    MMF_CUDA(cudaStreamSynchronize(p->streams[1]));
    MMF_CUDA(cudaFree(p->host_ws));
    p->host_ws = nullptr;
    p->host_ws_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&p->host_ws, bytes);
  if (e != cudaSuccess) return fail(MMF_ERR_NOMEM, std::string("host-path workspace cudaMalloc: ") + cudaGetErrorString(e));
  p->host_ws_bytes = bytes;
  return MMF_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// carve `bytes` out of a bump pointer
static void* carve(unsigned char*& cur, size_t bytes) {
  void* r = cur;
  cur += align_up(bytes, 256);
  return r;
}

static int fill_sos(const double* sos, int n_sections, SosArgs* a) {
  if (n_sections < 1 || n_sections > 16) return fail(MMF_ERR_UNSUPPORTED, "n_sections must be in [1, 16]");
  std::memset(a, 0, sizeof(*a));  // (also the padding: the struct is used as a cache key)
  a->n_sections = n_sections;
  std::memset(a->sos, 0, sizeof(a->sos));
  std::memset(a->zi, 0, sizeof(a->zi));
  for (int s = 0; s < n_sections; ++s)
    for (int k = 0; k < 6; ++k) a->sos[s][k] = sos[6 * s + k] / (k == 3 ? 1.0 : sos[6 * s + 3]);
  double zi[32];
  int padlen = 0;
  if (host_sos_zi(sos, n_sections, zi, &padlen) != 0) return fail(MMF_ERR_INVALID, "singular SOS section");
  for (int s = 0; s < n_sections; ++s) {
    a->zi[s][0] = zi[2 * s];
    a->zi[s][1] = zi[2 * s + 1];
  }
  a->padlen = padlen;
  return MMF_OK;
}

}  // namespace mmf


namespace mmf {

// Tables of the trajectory FFT live in the plan and are rebuilt only when the
// window / FFT size / band edges change.
static int ensure_mod_tables(mmf_plan* p, int win, int nfft, const int* lo, const int* hi, int n_bands) {
  bool same = p->mod_win == win && p->mod_nfft == nfft && p->mod_n_bands == n_bands;
  for (int b = 0; same && b < n_bands; ++b) same = p->mod_lo[b] == lo[b] && p->mod_hi[b] == hi[b];
  if (same) return MMF_OK;
  MMF_CUDA(cudaDeviceSynchronize());
  cudaFree(p->d_mod_hann);
  cudaFree(p->d_mod_tw1);
  cudaFree(p->d_mod_tw2);
  cudaFree(p->d_mod_lo);
  cudaFree(p->d_mod_hi);
  cudaFree(p->d_mod_g);
  p->d_mod_g = nullptr;
  p->mod_g_kp = 0;
  p->d_mod_hann = nullptr;
  p->d_mod_tw1 = p->d_mod_tw2 = nullptr;
  p->d_mod_lo = p->d_mod_hi = nullptr;
  p->mod_n_bands = -1;
  std::vector<float> hann(nfft, 0.0f);
  for (int i = 0; i < win; ++i) hann[i] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)i / (double)win));
  StftGeometry g;
  modspec_geometry(nfft, &g);
  std::vector<float2> tw1, tw2;
  if (modspec_fast_supported(nfft)) host_twiddles(nfft, g, tw1, tw2);
  std::vector<int> vlo(16, 0), vhi(16, 0);
  for (int b = 0; b < n_bands; ++b) {
    vlo[b] = lo[b];
    vhi[b] = hi[b];
  }
  cudaError_t e;
  if ((e = upload(&p->d_mod_hann, hann)) != cudaSuccess || (e = upload(&p->d_mod_tw1, tw1)) != cudaSuccess ||
      (e = upload(&p->d_mod_tw2, tw2)) != cudaSuccess || (e = upload(&p->d_mod_lo, vlo)) != cudaSuccess ||
      (e = upload(&p->d_mod_hi, vhi)) != cudaSuccess)
    return cuda_fail(e, "uploading modulation-spectrum tables");
  if (modspec_tc_supported(win, nfft) && !(p->cfg.flags & MMF_FLAG_NO_TC_MODSPEC)) {
    std::vector<uint16_t> g;
    uint16_t* d_g = nullptr;
    modspec_tc_table(win, nfft, g, &p->mod_g_kp);
    if ((e = upload(&d_g, g)) != cudaSuccess) return cuda_fail(e, "uploading modulation-spectrum GEMM operand");
    p->d_mod_g = d_g;
  }
  p->mod_win = win;
  p->mod_nfft = nfft;
  p->mod_n_bands = n_bands;
  for (int b = 0; b < 16; ++b) {
    p->mod_lo[b] = vlo[b];
    p->mod_hi[b] = vhi[b];
  }
  return MMF_OK;
}

static int run_modspec(mmf_plan* p, const float* mfcc, int64_t n_clips, int n_coef, int64_t T, int win, int hop,
                       int nfft, float* mag, float* band, const int* lo, const int* hi, int n_bands, cudaStream_t st) {
  if (T < win) return MMF_OK;
  if (!band) n_bands = 0;
  int rc = ensure_mod_tables(p, win, nfft, lo, hi, n_bands);
  if (rc) return rc;
  cudaError_t e;
  if (p->d_mod_g != nullptr) {
    // windowed DFT of 128 trajectory windows at a time as one tcgen05 GEMM (modspec_tc.cu)
    bool handled = false;
    e = modspec_tc_launch(mfcc, n_clips, n_coef, T, win, hop, nfft, p->d_mod_g, p->mod_g_kp, mag,
                          n_bands > 0 ? band : nullptr, p->d_mod_lo, p->d_mod_hi, n_bands, p->sm_count, st, &handled);
    if (e != cudaSuccess) return cuda_fail(e, "modspec_tc_kernel launch");
    if (handled) return MMF_OK;
  }
  if (modspec_fast_supported(nfft)) {
    e = modspec_fast_launch(mfcc, n_clips, n_coef, T, win, hop, nfft, p->d_mod_hann, p->d_mod_tw1, p->d_mod_tw2, mag,
                            n_bands > 0 ? band : nullptr, p->d_mod_lo, p->d_mod_hi, n_bands, p->sm_count, st);
  } else {
    e = modspec_launch(mfcc, n_clips, n_coef, T, win, hop, nfft, mag, n_bands > 0 ? band : nullptr, p->d_mod_lo,
                       p->d_mod_hi, n_bands, st);
  }
  if (e != cudaSuccess) return cuda_fail(e, "modspec kernel launch");
  return MMF_OK;
}

}  // namespace mmf

using namespace mmf;

extern "C" {

int mmf_version(void) { return MMF_VERSION; }
int mmf_abi_sizeof(int32_t which) {
  switch (which) {
    case 0: return (int)sizeof(mmf_config);
    case 1: return (int)sizeof(mmf_change_params);
    case 2: return (int)sizeof(mmf_modspec_params);
    default: return -1;
  }
}
const char* mmf_last_error(void) { return g_err.c_str(); }
int64_t mmf_launch_count(int32_t reset) {
  const long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int64_t mmf_num_frames(int64_t n_samples, int32_t n_fft, int32_t hop_length) {
  if (hop_length < 1) return -1;
  const int64_t padded = n_samples + 2 * (int64_t)(n_fft / 2);
  if (padded < n_fft) return 0;
  return 1 + (padded - n_fft) / hop_length;
}

int mmf_host_tables(const mmf_config* cfg, float* window, float* mel, float* dct) {
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  const int F = cfg->n_fft / 2 + 1;
  if (window) {
    std::vector<float> w;
    host_window(cfg->win_length, cfg->n_fft, w);
    std::memcpy(window, w.data(), w.size() * sizeof(float));
  }
  if (mel) {
    std::vector<float> m;
    std::vector<double> mf;
    host_mel_dense(cfg->sample_rate, cfg->n_fft, cfg->n_mels, cfg->fmin, cfg->fmax, m, mf);
    std::memcpy(mel, m.data(), (size_t)cfg->n_mels * F * sizeof(float));
  }
  if (dct) {
    std::vector<float> d;
    host_dct(cfg->n_mfcc, cfg->n_mels, d);
    std::memcpy(dct, d.data(), d.size() * sizeof(float));
  }
  return MMF_OK;
}

int mmf_sos_zi(const double* sos, int32_t n_sections, double* zi, int32_t* padlen) {
  if (!sos || !zi) return fail(MMF_ERR_INVALID, "NULL argument");
  int pl = 0;
  if (host_sos_zi(sos, n_sections, zi, &pl) != 0) return fail(MMF_ERR_INVALID, "bad SOS cascade");
  if (padlen) *padlen = pl;
  return MMF_OK;
}

int mmf_plan_create(mmf_plan** out, const mmf_config* cfg) {
  if (!out) return fail(MMF_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(MMF_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                  cudaGetErrorString(ce));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(MMF_ERR_INVALID, "device ordinal out of range");
  MMF_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  MMF_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10)
    return fail(MMF_ERR_UNSUPPORTED, "this library is built for sm_100a (Blackwell B200) only");

  {
    // stream-ordered scratch (cudaMallocAsync in the resampler, the Hilbert envelope and the generic-n_fft path) stays
    // cached in the device's default pool instead of going back to the driver at every synchronisation: with the
    // default release threshold of 0 the Hilbert envelope of one 10 s clip took 1.1 .. 118 ms from call to call
    static bool pool_set[64] = {};
    if (cfg->device < 64 && !pool_set[cfg->device]) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, cfg->device) == cudaSuccess) {
        uint64_t keep = 1ull << 30;  // up to 1 GiB of freed scratch is kept for reuse
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_set[cfg->device] = true;
    }
  }
  mmf_plan* p = new mmf_plan();
  p->cfg = *cfg;
  p->F = cfg->n_fft / 2 + 1;
  p->sm_count = prop.multiProcessorCount;
  // the register FFT covers powers of two in [256, 4096]; every other n_fft librosa accepts goes through the
  // FP32 matrix-product DFT (dft_generic.cu)
  p->generic = (cfg->n_fft < 256 || (cfg->n_fft & (cfg->n_fft - 1))) ? 1 : 0;
  p->packed = (!p->generic && stft_packed_supported(cfg->n_fft) && !(cfg->flags & MMF_FLAG_SCALAR_FFT)) ? 1 : 0;
  p->geo = StftGeometry{};
  if (!p->generic) stft_geometry(cfg->n_fft, p->packed, 256, &p->geo);
  p->lead = cfg->preemph != 0.0f ? 2 : 0;

  // ---- constant tables
  std::vector<float> window, mel, dct;
  std::vector<double> mel_f;
  host_window(cfg->win_length, cfg->n_fft, window);
  host_mel_dense(cfg->sample_rate, cfg->n_fft, cfg->n_mels, cfg->fmin, cfg->fmax, mel, mel_f);
  MelSparse sp;
  if (!host_mel_sparse(mel, mel_f, cfg->sample_rate, cfg->n_fft, cfg->n_mels, sp)) {
    delete p;
    return fail(MMF_ERR_UNSUPPORTED, "mel filterbank is not a two-slope (triangular) bank");
  }
  host_dct(cfg->n_mfcc, cfg->n_mels, dct);
  if (p->generic) {
    std::vector<float> tab;
    dft_generic_table(cfg->n_fft, window, tab, &p->dft_ld);
    p->nc_pad = cfg->n_mfcc <= 16 ? 16 : (cfg->n_mfcc <= 32 ? 32 : (cfg->n_mfcc <= 64 ? 64 : 128));
    std::vector<float> dct_pad((size_t)cfg->n_mels * p->nc_pad, 0.0f);
    for (int k = 0; k < cfg->n_mfcc; ++k)
      for (int m = 0; m < cfg->n_mels; ++m) dct_pad[(size_t)m * p->nc_pad + k] = dct[(size_t)k * cfg->n_mels + m];
    std::vector<float2> w2(p->F);
    for (int k = 0; k < p->F; ++k) w2[k] = make_float2(sp.w2[2 * k], sp.w2[2 * k + 1]);
    std::vector<float4> dct_bfrag;
    mfcc_mma_bfrag(dct.data(), cfg->n_mfcc, cfg->n_mels, dct_bfrag);
    cudaError_t e;
    if ((e = upload(&p->d_dft_tab, tab)) != cudaSuccess || (e = upload(&p->d_window, window)) != cudaSuccess ||
        (e = upload(&p->d_seg, sp.seg_start)) != cudaSuccess || (e = upload(&p->d_w2, w2)) != cudaSuccess ||
        (e = upload(&p->d_dct, dct_pad)) != cudaSuccess || (e = upload(&p->d_dct_bfrag, dct_bfrag)) != cudaSuccess) {
      mmf_plan_destroy(p);
      return cuda_fail(e, "uploading plan constants (generic n_fft)");
    }
    for (int i = 0; i < 2; ++i) cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking);
    for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&p->events[i], cudaEventDisableTiming);
    *out = p;
    return MMF_OK;
  }
  // grouped mel walk on the bin-pair power tile: the default; MMF_FLAG_MEL_WALK / MMF_FLAG_MMA_MEL select
  // the [bin][frame] tile with the sparse walk / the mma.sync projection
  MelGroups mg;
  host_mel_groups(sp, p->F, cfg->n_mels, mg);
  p->mel_groups = (cfg->flags & (MMF_FLAG_MEL_WALK | MMF_FLAG_MMA_MEL)) ? 0 : 1;
  p->mg_n = mg.n_groups;

  // ---- tensor-core mel tables.  An n-tile is a run of up to 8 consecutive bands (the N of
  // mma.m16n8k8); it is cut short when its bands span more than kTileBlocks k-tiles of 8 bins, so
  // that the wide top bands do not pile all their blocks on one warp.  Per tile: the k-tiles whose
  // 8x8 block of the filterbank is not all zero, and per block the B fragments
  // (b0: k = lane%4, b1: k = lane%4 + 4; n = lane/4) as plain fp32 (split into TF32 hi/lo on the fly).
  std::vector<float2> mma_bw;
  std::vector<int> mma_pk8, mma_npair(1, 0), mma_tile;  // mma_tile[j] = first band | bands << 16
  {
    const int F = p->F, kTileBlocks = 8;
    auto wgt = [&](int k, int band) -> float {
      return (k < F && band < cfg->n_mels) ? mel[(size_t)band * F + k] : 0.0f;
    };
    auto blocks_of = [&](int b0, int nb, std::vector<int>* out) {
      int cnt = 0;
      for (int k8 = 0; k8 < (F + 7) / 8; ++k8) {
        bool any = false;
        for (int k = 8 * k8; k < 8 * k8 + 8 && !any; ++k)
          for (int b = b0; b < b0 + nb && !any; ++b) any = wgt(k, b) != 0.0f;
        if (any) {
          ++cnt;
          if (out) out->push_back(k8);
        }
      }
      return cnt;
    };
    for (int b0 = 0; b0 < cfg->n_mels;) {
      int nb = 1;
      while (nb < 8 && b0 + nb < cfg->n_mels && blocks_of(b0, nb + 1, nullptr) <= kTileBlocks) ++nb;
      std::vector<int> ks;
      blocks_of(b0, nb, &ks);
      for (int k8 : ks) {
        mma_pk8.push_back(k8);
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t4 = lane & 3;
          const bool own = g < nb;  // columns beyond the tile's bands belong to the next tile
          mma_bw.push_back(make_float2(own ? wgt(8 * k8 + t4, b0 + g) : 0.0f, own ? wgt(8 * k8 + t4 + 4, b0 + g) : 0.0f));
        }
      }
      mma_tile.push_back(b0 | (nb << 16));
      mma_npair.push_back((int)mma_pk8.size());
      b0 += nb;
    }
  }
  const int NT = (int)mma_tile.size();
  p->mma_n_pairs = (int)mma_pk8.size();

  // ---- tile geometry.  Preference: the widest tile first (32 frames = one frame per lane /
  // two MMA row tiles in the mel phase), then as much double buffering as two CTAs per SM allow
  // (<= 113 KB each: 227 KB per SM, 1 KB reserved per CTA); one CTA per SM as the last resort.
  const int fpw = p->geo.tpf < 32 ? 32 / p->geo.tpf : 1;  // frames per warp in the FFT phase
  auto pitch_for = [&](int tf) {
    // bin-pair layout: 8-byte words per bin-pair row; = 2 (mod 16) keeps the 32-bit stores of the two
    // thread groups of a warp on disjoint banks (stft_core.cuh: TilePairs)
    if (p->mel_groups) return std::max(tf, 2) + 2;
    if (p->packed) {
      // two-frame path: 64-bit stores of (frame, frame+1) pairs by consecutive bins:
      // pitch == 2 (mod 4) keeps them 8-byte aligned and conflict-free per half-warp
      int pp = std::max(tf, 2);
      while (pp % 4 != 2) ++pp;
      return pp;
    }
    int pp = std::max(tf, fpw);
    if (fpw == 1) {
      if ((pp & 1) == 0) ++pp;
    } else {
      pp = (pp + fpw - 1) / fpw * fpw;
      if (((pp / fpw) & 1) == 0) pp += fpw;
    }
    return pp;
  };
  const size_t budget2 = 113 * 1024, budget1 = 226 * 1024;
  static const int kBufChoices[4][2] = {{2, 2}, {1, 2}, {2, 1}, {1, 1}};  // {span_bufs, pt_bufs}
  struct Geo {
    int tf = 0, span_bufs = 2, pt_bufs = 2, ctas = 2, threads = 256;
    size_t smem = 0;
  };
  auto smem_for = [&](int tf, int span_bufs, int pt_bufs, int mel_mma, int threads) {
    const int span = (tf - 1) * cfg->hop_length + cfg->n_fft + p->lead + 3;
    const int alloc = (span + 255) / 256 * 256;
    return stft_smem_bytes(cfg->n_fft, alloc, span_bufs, pitch_for(tf), pt_bufs, p->packed,
                           stft_mel_table_bytes(cfg->n_fft, cfg->n_mels, mel_mma, p->mma_n_pairs, NT, p->mel_groups, p->mg_n),
                           threads, p->mel_groups);
  };
  auto fpi_for = [&](int threads) { return (threads / p->geo.tpf) * (p->packed ? 2 : 1); };
  // widest tile (then most double buffering) that fits `budget` with CTAs of `threads`
  auto fit = [&](int mel_mma, int threads, size_t budget, Geo* g) {
    const int fpi = fpi_for(threads);
    if (fpi < 1 || fpi > 32) return false;
    for (int cand = 32; cand >= 1; cand >>= 1) {
      if (cand < fpi || cand % fpi) continue;
      for (const auto& ch : kBufChoices) {
        const size_t b = smem_for(cand, ch[0], ch[1], mel_mma, threads);
        if (b <= budget) {
          g->tf = cand;
          g->span_bufs = ch[0];
          g->pt_bufs = ch[1];
          g->smem = b;
          g->threads = threads;
          return true;
        }
      }
    }
    return false;
  };
  auto choose = [&](int mel_mma) {
    Geo g;
    g.ctas = 2;
    if (fit(mel_mma, 256, budget2, &g)) return g;
    g.ctas = 1;
    // one CTA per SM: 512 threads (16 warps) when the transform groups are whole warps and the
    // tiles still fit, else 256
    Geo g512;
    g512.ctas = 1;
    const bool ok512 = cfg->n_fft >= 1024 && !mel_mma && fit(mel_mma, 512, budget1, &g512);
    const bool ok256 = fit(mel_mma, 256, budget1, &g);
    if (ok512 && (!ok256 || g512.tf >= g.tf) && !std::getenv("MMF_NO_512")) return g512;
    if (!ok256) g.tf = 0;
    return g;
  };
  // The tensor-core mel (MMF_FLAG_MMA_MEL) keeps its B fragments in shared memory: honoured unless
  // that costs tile width or the second CTA per SM (very wide filterbanks).  Default is the sparse
  // FP32 walk: the filterbank is 95-98 % zeros and the split-precision MMA path measured 3-7 %
  // slower on every BASELINE shape.
  Geo gs = choose(0), gm = choose(1);
  const bool units_fit = NT * 2 <= 8 * 16 && NT <= 255;
  p->mel_mma = ((cfg->flags & MMF_FLAG_MMA_MEL) && units_fit && gm.tf != 0 && gm.tf >= gs.tf &&
                gm.ctas >= gs.ctas)
                   ? 1
                   : 0;
  Geo g = p->mel_mma ? gm : gs;
  // debugging / tuning overrides
  if (const char* e = std::getenv("MMF_TF")) {
    const int v = std::atoi(e);
    const int fpi = fpi_for(256);
    g.threads = 256;
    if (v >= fpi && v <= 32 && (v & (v - 1)) == 0 && v % fpi == 0) g.tf = v;
    if (const char* e2 = std::getenv("MMF_SPAN_BUFS")) g.span_bufs = std::atoi(e2) == 1 ? 1 : 2;
    if (const char* e3 = std::getenv("MMF_PT_BUFS")) g.pt_bufs = std::atoi(e3) == 1 ? 1 : 2;
    g.smem = smem_for(g.tf, g.span_bufs, g.pt_bufs, p->mel_mma, 256);
    g.ctas = g.smem <= budget2 ? 2 : 1;
    if (g.smem > budget1) g.tf = 0;
  }
  if (g.tf == 0) {
    delete p;
    return fail(MMF_ERR_UNSUPPORTED, "hop_length/n_fft combination needs more shared memory than one SM has");
  }
  p->threads = g.threads;
  stft_geometry(cfg->n_fft, p->packed, p->threads, &p->geo);
  const int tf = g.tf;
  p->TF = tf;
  p->pt_bufs = g.pt_bufs;
  p->span_bufs = g.span_bufs;
  p->ctas_per_sm = g.ctas;
  p->ppitch = pitch_for(tf);
  p->smem = g.smem;
  p->mel_tab_bytes = stft_mel_table_bytes(cfg->n_fft, cfg->n_mels, p->mel_mma, p->mma_n_pairs, NT, p->mel_groups, p->mg_n);
  p->mma_n_tiles = NT;

  // ---- tensor-core mel work list: unit = (n-tile, 16-frame m-tile), cost = blocks of the n-tile;
  // longest-processing-time assignment to the 8 warps (each unit is produced by exactly one warp)
  std::vector<int> mma_units(8 * 16, -1);
  if (p->mel_mma) {
    const int MT = (tf + 15) / 16;
    std::vector<std::pair<int, int>> units;  // (cost, n | m << 8)
    for (int n = 0; n < NT; ++n)
      for (int m = 0; m < MT; ++m) units.push_back({mma_npair[n + 1] - mma_npair[n], n | (m << 8)});
    std::stable_sort(units.begin(), units.end(),
                     [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first > b.first; });
    long load[8] = {0};
    int cnt[8] = {0};
    for (const auto& u : units) {
      int w = 0;
      for (int i = 1; i < 8; ++i)
        if (load[i] < load[w] || (load[i] == load[w] && cnt[i] < cnt[w])) w = i;
      if (cnt[w] >= 16) {  // cannot happen with units_fit, kept as a guard
        delete p;
        return fail(MMF_ERR_UNSUPPORTED, "too many mel tiles per warp");
      }
      mma_units[w * 16 + cnt[w]++] = u.second;
      load[w] += u.first + 2;  // + epilogue
    }
  }

  // ---- mel band groups of the sparse FP32 walk: 8 warps * (32 / TF) workers, contiguous bands
  // each, balanced on cost = bins walked + bands emitted (smallest achievable maximum, greedy
  // fill under a binary-searched bound)
  std::vector<int> band_split(257, cfg->n_mels);
  {
    const int workers = std::min(256, (p->threads / 32) * (32 / tf));
    const int cbin = 7, cband = 48;  // measured: 7 instructions per bin, 29 per segment + 19 per band
    const int cgroup = 14, cseg = 41;  // grouped walk (SASS count): per 4-bin weight group, per segment
    auto group_cost = [&](int a, int b) -> long {  // bands [a, b): segments a..b
      if (p->mel_groups) return (long)cgroup * (mg.segtab[2 * (b + 1) + 1] - mg.segtab[2 * a + 1]) + (long)cseg * (b - a + 1);
      return cbin * (sp.seg_start[b + 1] - sp.seg_start[a]) + cband * (b - a);
    };
    auto fill = [&](long bound, std::vector<int>* out) {
      int a = 0, used = 0;
      while (a < cfg->n_mels) {
        if (used == workers) return false;
        int b = a + 1;
        if (group_cost(a, b) > bound) return false;
        while (b < cfg->n_mels && group_cost(a, b + 1) <= bound) ++b;
        if (out) (*out)[used] = a;
        a = b;
        ++used;
      }
      if (out)
        for (int w = used; w < 257; ++w) (*out)[w] = cfg->n_mels;
      return true;
    };
    long lo = 0, hi = group_cost(0, cfg->n_mels);
    while (lo < hi) {
      const long mid = (lo + hi) / 2;
      if (fill(mid, nullptr)) hi = mid; else lo = mid + 1;
    }
    fill(lo, &band_split);
  }
  std::vector<float2> tw1, tw2;
  host_twiddles(cfg->n_fft, p->geo, tw1, tw2);
  p->nc_pad = cfg->n_mfcc <= 16 ? 16 : (cfg->n_mfcc <= 32 ? 32 : (cfg->n_mfcc <= 64 ? 64 : 128));
  std::vector<float> dct_pad((size_t)cfg->n_mels * p->nc_pad, 0.0f);
  for (int k = 0; k < cfg->n_mfcc; ++k)
    for (int m = 0; m < cfg->n_mels; ++m) dct_pad[(size_t)m * p->nc_pad + k] = dct[(size_t)k * cfg->n_mels + m];
  std::vector<float2> w2(p->F);
  for (int k = 0; k < p->F; ++k) w2[k] = make_float2(sp.w2[2 * k], sp.w2[2 * k + 1]);

  std::vector<float4> dct_bfrag;
  mfcc_mma_bfrag(dct.data(), cfg->n_mfcc, cfg->n_mels, dct_bfrag);
  cudaError_t e;
  {
    std::vector<int2> segtab(cfg->n_mels + 2);
    for (int j = 0; j < cfg->n_mels + 2; ++j) segtab[j] = make_int2(mg.segtab[2 * j], mg.segtab[2 * j + 1]);
    std::vector<float4> w4((size_t)2 * mg.n_groups);
    for (size_t i = 0; i < w4.size(); ++i) w4[i] = make_float4(mg.w[4 * i], mg.w[4 * i + 1], mg.w[4 * i + 2], mg.w[4 * i + 3]);
    std::vector<int2> segstep(cfg->n_mels + 3);
    for (int j = 0; j < cfg->n_mels + 3; ++j)  // step in 8-byte tile words: two bin-pair rows per group
      segstep[j] = make_int2(mg.segstep[2 * j], mg.segstep[2 * j + 1] * 2 * p->ppitch);
    if ((e = upload(&p->d_mg_seg, segtab)) != cudaSuccess || (e = upload(&p->d_mg_w, w4)) != cudaSuccess ||
        (e = upload(&p->d_mg_step, segstep)) != cudaSuccess) {
      mmf_plan_destroy(p);
      return cuda_fail(e, "uploading the grouped mel tables");
    }
  }
  if ((e = upload(&p->d_dct_bfrag, dct_bfrag)) != cudaSuccess || (e = upload(&p->d_window, window)) != cudaSuccess || (e = upload(&p->d_tw1, tw1)) != cudaSuccess ||
      (e = upload(&p->d_tw2, tw2)) != cudaSuccess || (e = upload(&p->d_seg, sp.seg_start)) != cudaSuccess ||
      (e = upload(&p->d_band_split, band_split)) != cudaSuccess ||
      (e = upload(&p->d_mma_bw, mma_bw)) != cudaSuccess || (e = upload(&p->d_mma_pk8, mma_pk8)) != cudaSuccess ||
      (e = upload(&p->d_mma_npair, mma_npair)) != cudaSuccess ||
      (e = upload(&p->d_mma_tile, mma_tile)) != cudaSuccess ||
      (e = upload(&p->d_mma_units, mma_units)) != cudaSuccess ||
      (e = upload(&p->d_w2, w2)) != cudaSuccess || (e = upload(&p->d_dct, dct_pad)) != cudaSuccess) {
    mmf_plan_destroy(p);
    return cuda_fail(e, "uploading plan constants");
  }

  // mel projection as a tcgen05 GEMM (default where it applies; stft_mel_tc.cu)
  if (!(cfg->flags & (MMF_FLAG_NO_TC_MEL | MMF_FLAG_MEL_WALK | MMF_FLAG_MMA_MEL | MMF_FLAG_SPLIT_SMEM)) &&
      stft_mel_tc_supported(cfg->n_fft, stft_mel_tc_active_bands(mel, cfg->n_mels, p->F), cfg->hop_length, p->lead,
                            p->packed)) {
    std::vector<uint16_t> tab;
    uint16_t* d_t = nullptr;
    stft_mel_tc_table(mel, cfg->n_mels, p->F, tab);
    if ((e = upload(&d_t, tab)) != cudaSuccess) {
      mmf_plan_destroy(p);
      return cuda_fail(e, "uploading the tensor-core mel operand");
    }
    p->d_mel_tc = d_t;
    p->mel_tc_act = stft_mel_tc_active_bands(mel, cfg->n_mels, p->F);
    // thread tau of a frame group holds samples 2 (tau + 16 n2), + 1: n2 selects a run of 32 samples
    p->win_lo = 16;
    p->win_hi = 0;
    for (int n2 = 0; n2 < 16; ++n2)
      for (int i = 32 * n2; i < 32 * n2 + 32; ++i)
        if (window[i] != 0.0f) {
          p->win_lo = std::min(p->win_lo, n2);
          p->win_hi = std::max(p->win_hi, n2 + 1);
        }
    if (p->win_hi <= p->win_lo) p->win_lo = 0, p->win_hi = 16;
  }
  if ((cfg->flags & MMF_FLAG_TC_DCT) && !(cfg->flags & MMF_FLAG_MMA_DCT) && mfcc_tc_supported(cfg->n_mfcc, cfg->n_mels)) {
    std::vector<uint16_t> tab;
    uint16_t* d_t = nullptr;
    mfcc_tc_table(dct.data(), cfg->n_mfcc, cfg->n_mels, tab, &p->dct_tc_kp);
    if ((e = upload(&d_t, tab)) != cudaSuccess) {
      mmf_plan_destroy(p);
      return cuda_fail(e, "uploading the tensor-core DCT operand");
    }
    p->d_dct_tc = d_t;
  }
  if ((cfg->flags & MMF_FLAG_TC_FFT) && tc_fft_supported(cfg->n_fft, cfg->hop_length, cfg->preemph)) {
    std::vector<uint16_t> btab;
    std::vector<float> twf;
    tc_fft_tables(btab, twf);
    uint16_t* d_b = nullptr;
    float* d_t = nullptr;
    if ((e = upload(&d_b, btab)) != cudaSuccess || (e = upload(&d_t, twf)) != cudaSuccess) {
      cudaFree(d_b);
      mmf_plan_destroy(p);
      return cuda_fail(e, "uploading tensor-core transform tables");
    }
    p->d_tc_btab = d_b;
    p->d_tc_tw = reinterpret_cast<float2*>(d_t);
  }

  // ---- driver entry point for tensor-map encoding (no link-time libcuda dependency)
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) p->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  for (int i = 0; i < 2; ++i) cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking);
  for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&p->events[i], cudaEventDisableTiming);
  *out = p;
  return MMF_OK;
}

int mmf_plan_destroy(mmf_plan* p) {
  if (!p) return MMF_OK;
  cudaSetDevice(p->cfg.device);
  cudaDeviceSynchronize();
  cudaFree(p->d_window);
  cudaFree(p->d_dft_tab);
  cudaFree(p->d_dct_tc);
  cudaFree(p->d_tc_btab);
  cudaFree(p->d_tc_tw);
  cudaFree(p->d_tw1);
  cudaFree(p->d_tw2);
  cudaFree(p->d_seg);
  cudaFree(p->d_band_split);
  cudaFree(p->d_mma_bw);
  cudaFree(p->d_mma_pk8);
  cudaFree(p->d_mma_npair);
  cudaFree(p->d_mma_tile);
  cudaFree(p->d_mma_units);
  cudaFree(p->d_w2);
  cudaFree(p->d_mg_seg);
  cudaFree(p->d_mg_step);
  cudaFree(p->d_mg_w);
  cudaFree(p->d_dct);
  cudaFree(p->d_mel_tc);
  cudaFree(p->d_dct_bfrag);
  cudaFree(p->ws);
  cudaFree(p->host_ws);
  cudaFree(p->d_mod_hann);
  cudaFree(p->d_mod_tw1);
  cudaFree(p->d_mod_tw2);
  cudaFree(p->d_mod_lo);
  cudaFree(p->d_mod_hi);
  cudaFree(p->d_mod_g);
  if (p->pinned) cudaFreeHost(p->pinned);
  for (int i = 0; i < 2; ++i)
    if (p->streams[i]) cudaStreamDestroy(p->streams[i]);
  for (int i = 0; i < 4; ++i)
    if (p->events[i]) cudaEventDestroy(p->events[i]);
  delete p;
  return MMF_OK;
}

}  // extern "C"

namespace mmf {

// shared body of mmf_stft_power / mmf_logmel
static int run_stft(mmf_plan* p, const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                    float* power, float* logmel, int* clipmax, cudaStream_t st) {
  if (!p) return fail(MMF_ERR_INVALID, "plan is NULL");
  if (!pcm) return fail(MMF_ERR_INVALID, "pcm is NULL");
  if (n_clips < 1 || n_samples < 1) return fail(MMF_ERR_INVALID, "n_clips and n_samples must be positive");
  if (clip_stride < n_samples && n_clips > 1) return fail(MMF_ERR_INVALID, "clip_stride < n_samples");
  if (logmel && !clipmax) return fail(MMF_ERR_INVALID, "clipmax buffer required with logmel");
  const mmf_config& c = p->cfg;
  const int64_t T = mmf_num_frames(n_samples, c.n_fft, c.hop_length);
  if (T < 1) return fail(MMF_ERR_INVALID, "input too short for one frame");
  if (T > 0x7fffffff || n_samples + c.n_fft > 0x7fffffffLL)
    return fail(MMF_ERR_UNSUPPORTED, "clip longer than 2^31 samples");
  MMF_CUDA(cudaSetDevice(c.device));

  if (p->generic) {
    // FP32 matrix-product DFT -> power spectrum in HBM (the caller's buffer, or stream-ordered scratch in clip
    // chunks of <= 1 GiB) -> sparse mel walk + log, one frame per thread
    if (clipmax) {
      cudaError_t e = fill_i32_launch(clipmax, n_clips, (int)0x80800000, st);
      if (e != cudaSuccess) return cuda_fail(e, "clipmax init");
    }
    const size_t per_clip = (size_t)p->F * T * 4;
    const int64_t chunk = power ? n_clips : std::max<int64_t>(1, std::min<int64_t>(n_clips, ((size_t)1 << 30) / per_clip));
    float* scratch = nullptr;
    if (!power) MMF_CUDA(cudaMallocAsync((void**)&scratch, (size_t)chunk * per_clip, st));
    cudaError_t e = cudaSuccess;
    for (int64_t c0 = 0; c0 < n_clips && e == cudaSuccess; c0 += chunk) {
      const int64_t nc = std::min<int64_t>(chunk, n_clips - c0);
      float* pw = power ? power + (size_t)c0 * p->F * T : scratch;
      e = dft_generic_power_launch(pcm + (size_t)c0 * clip_stride, nc, n_samples, clip_stride, (int)T, c.hop_length,
                                   c.n_fft, c.preemph, p->d_dft_tab, p->dft_ld, pw, st);
      if (e == cudaSuccess && logmel)
        e = mel_from_power_launch(pw, nc, p->F, (int)T, c.n_mels, c.amin, p->d_seg, p->d_w2,
                                  logmel + (size_t)c0 * c.n_mels * T, clipmax + c0, st);
    }
    if (scratch) cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return cuda_fail(e, "generic DFT kernels launch");
    return MMF_OK;
  }

  StftArgs a{};
  a.pcm = pcm;
  a.n_samples = n_samples;
  a.clip_stride = clip_stride;
  a.T = (int)T;
  a.hop = c.hop_length;
  a.TF = p->TF;
  a.tiles_per_clip = (int)((T + p->TF - 1) / p->TF);
  a.n_tiles = (long)a.tiles_per_clip * n_clips;
  a.lead = p->lead;
  a.span_floats = (p->TF - 1) * c.hop_length + c.n_fft + p->lead + 3;  // +3: 16-byte aligned start
  a.span_alloc = (a.span_floats + 255) / 256 * 256;
  a.ppitch = p->ppitch;
  a.pt_bufs = p->pt_bufs;
  a.packed = p->packed;
  a.threads = p->threads;
  a.mel_mma = p->mel_mma;
  a.m_tiles = (p->TF + 15) / 16;
  a.mma_n_pairs = p->mma_n_pairs;
  a.mel_tab_bytes = (int)p->mel_tab_bytes;
  a.mma_bw = p->d_mma_bw;
  a.mma_pk8 = p->d_mma_pk8;
  a.mma_npair = p->d_mma_npair;
  a.mma_tile = p->d_mma_tile;
  a.mma_n_tiles = p->mma_n_tiles;
  a.mma_units = p->d_mma_units;
  a.early_tma = 1;
#ifdef MMF_PROFILE  // profiling builds only: a stray environment variable must never change results
  if (const char* e = std::getenv("MMF_DEBUG_SKIP")) a.debug_skip = std::atoi(e);
  if (const char* e = std::getenv("MMF_EARLY_TMA")) a.early_tma = std::atoi(e);
#endif
  a.span_bufs = p->span_bufs;
  a.vec_ok = (c.hop_length % 2 == 0) ? 1 : 0;
  a.split_regs = (c.n_fft == 512 && !(c.flags & MMF_FLAG_SPLIT_SMEM)) ? 1 : 0;
  a.n_mels = c.n_mels;
  a.amin = c.amin;
  a.preemph = c.preemph;
  a.window = p->d_window;
  a.tw1 = p->d_tw1;
  a.tw2 = p->d_tw2;
  a.seg_start = p->d_seg;
  a.band_split = p->d_band_split;
  a.w2 = p->d_w2;
  a.mel_groups = p->mel_groups;
  a.mg_n = p->mg_n;
  a.mg_seg = p->d_mg_seg;
  a.mg_step = p->d_mg_step;
  a.mg_w = p->d_mg_w;
  a.logmel = logmel;
  a.clipmax = clipmax;
  a.power = power;

  // TMA needs a 16-byte aligned base and row pitch; otherwise use the plain loader
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  bool tma = p->encode != nullptr && !(c.flags & MMF_FLAG_NO_TMA) && ((uintptr_t)pcm % 16 == 0) &&
             ((clip_stride * 4) % 16 == 0 || n_clips == 1);
  if (tma) {
    cuuint64_t gdim[2] = {(cuuint64_t)n_samples, (cuuint64_t)n_clips};
    cuuint64_t gstr[1] = {(cuuint64_t)(n_clips == 1 ? align_up((size_t)n_samples * 4, 16) : clip_stride * 4)};
    cuuint32_t box[2] = {256, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = p->encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)pcm, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) tma = false;
  }
  a.use_tma = tma ? 1 : 0;

  if (clipmax) {
    cudaError_t e = fill_i32_launch(clipmax, n_clips, (int)0x80800000, st)  /* key of -FLT_MAX */;
    if (e != cudaSuccess) return cuda_fail(e, "clipmax init");
  }
  // the mel projection on the tensor cores wherever the plan supports it -- a decision of the configuration alone, so
  // that a clip's features never depend on the size or the chunking of the batch it came in
  if (p->d_mel_tc && logmel && !power && ((T + 127) / 128) * n_clips < 0x7fffffffLL) {
    cudaError_t e = stft_mel_tc_launch(tmap, tma ? 1 : 0, pcm, n_clips, n_samples, clip_stride, (int)T, c.hop_length, p->lead, c.n_mels, p->mel_tc_act, c.amin,
                                       c.preemph, p->d_window, p->win_lo, p->win_hi, p->d_tw1, p->d_mel_tc, logmel, clipmax, p->sm_count, st);
    count_launch();
    if (e != cudaSuccess) return cuda_fail(e, "stft_mel_tc_kernel launch");
    return MMF_OK;
  }
  const long max_ctas = (long)p->ctas_per_sm * p->sm_count;
  const int grid = (int)std::min<long>(a.n_tiles, max_ctas);
  cudaError_t e = stft_mel_launch(c.n_fft, tmap, a, grid, p->smem, st);
  count_launch();
  if (e != cudaSuccess) return cuda_fail(e, "stft_mel_kernel launch");
  return MMF_OK;
}

// zero-phase IIR over independent rows: chunk-parallel kernel when the extended row fits in
// shared memory (<= 6112 samples, <= 4 sections); super-block scan for long rows; else the
// sequential kernel (many long rows, or more than 4 sections)
static cudaError_t sosfiltfilt_any(const void* x, int x_is_f32, long rows, long T, long xs, int group_rows,
                                   long group_stride, const SosArgs& a, double* y, long ys, cudaStream_t st) {
  SosPar par;
  if (sos_par_fill(a, T, &par))
    return sosfiltfilt_par_launch(x, x_is_f32, rows, T, xs, group_rows, group_stride, par, y, ys, st);
  // long rows, few of them: the sequential kernel would walk every sample one by one
  if (sos_long_supported(a, rows, T) && (size_t)rows * (size_t)(T + 2 * a.padlen) * 8 <= ((size_t)1 << 30))
    return sosfiltfilt_long_launch(x, x_is_f32, rows, T, xs, group_rows, group_stride, a, y, ys, st);
  return sosfiltfilt_launch_grouped(x, x_is_f32, rows, T, xs, group_rows, group_stride, a, y, ys, st);
}

}  // namespace mmf

extern "C" {

int mmf_stft_power(mmf_plan* plan, const float* pcm_dev, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                   float* power_dev, void* stream) {
  if (!power_dev) return fail(MMF_ERR_INVALID, "power_dev is NULL");
  if (plan && plan->d_tc_btab && pcm_dev && n_clips >= 1 && n_samples >= 1 && (clip_stride >= n_samples || n_clips == 1)) {
    // tcgen05 transform (MMF_FLAG_TC_FFT): fp16 x3 GEMM stages, accumulators in tensor memory
    const int64_t T = mmf_num_frames(n_samples, plan->cfg.n_fft, plan->cfg.hop_length);
    if (T >= 1 && T <= 0x7fffffff) {
      MMF_CUDA(cudaSetDevice(plan->cfg.device));
      cudaError_t e = tc_fft_power_launch(pcm_dev, n_clips, n_samples, clip_stride, (int)T, plan->cfg.hop_length,
                                          plan->d_window, plan->d_tc_btab, plan->d_tc_tw, power_dev, plan->sm_count,
                                          (cudaStream_t)stream);
      if (e != cudaSuccess) return cuda_fail(e, "tc_fft512_kernel launch");
      return MMF_OK;
    }
  }
  return run_stft(plan, pcm_dev, n_clips, n_samples, clip_stride, power_dev, nullptr, nullptr, (cudaStream_t)stream);
}

int mmf_logmel(mmf_plan* plan, const float* pcm_dev, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
               float* logmel_dev, int32_t* clipmax_dev, void* stream) {
  if (!logmel_dev) return fail(MMF_ERR_INVALID, "logmel_dev is NULL");
  return run_stft(plan, pcm_dev, n_clips, n_samples, clip_stride, nullptr, logmel_dev, clipmax_dev,
                  (cudaStream_t)stream);
}

int mmf_mfcc(mmf_plan* plan, float* logmel_dev, const int32_t* clipmax_dev, int64_t n_clips, int64_t T,
             float* mfcc_dev, float* delta_dev, int32_t clamp_in_place, void* stream) {
  if (!plan || !logmel_dev || !clipmax_dev || !mfcc_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (n_clips < 1 || T < 1) return fail(MMF_ERR_INVALID, "n_clips and T must be positive");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  for (int64_t c0 = 0; c0 < n_clips; c0 += 65535) {
    const int64_t nc = std::min<int64_t>(65535, n_clips - c0);
    float* lm = logmel_dev + (size_t)c0 * plan->cfg.n_mels * T;
    float* mf = mfcc_dev + (size_t)c0 * plan->cfg.n_mfcc * T;
    float* dl = delta_dev ? delta_dev + (size_t)c0 * plan->cfg.n_mfcc * T : nullptr;
    // Tensor-core DCT-II (mma.sync TF32 x3) on request only: measured 185 us against 95 us for the
    // FP32 kernel on the bench workload
    const bool mma = (plan->cfg.flags & MMF_FLAG_MMA_DCT) && mfcc_mma_supported(plan->cfg.n_mfcc, plan->cfg.n_mels);
    if (!mma && plan->d_dct_tc != nullptr) {
      // MMF_FLAG_TC_DCT: tcgen05 GEMM over 128-frame tiles (mfcc_tc.cu)
      cudaError_t et = mfcc_tc_launch(plan->d_dct_tc, plan->dct_tc_kp, lm, clipmax_dev + c0, nc, T, plan->cfg.n_mels,
                                      plan->cfg.n_mfcc, plan->cfg.top_db, mf, dl, clamp_in_place, plan->sm_count,
                                      (cudaStream_t)stream);
      if (et == cudaSuccess) continue;
      if (et != cudaErrorNotSupported) return cuda_fail(et, "mfcc_tc_kernel launch");
    }
    cudaError_t e = mma ? mfcc_mma_launch(plan->d_dct_bfrag, lm, clipmax_dev + c0, nc, T, plan->cfg.n_mels,
                                          plan->cfg.n_mfcc, plan->cfg.top_db, mf, dl, clamp_in_place, (cudaStream_t)stream)
                        : mfcc_launch(plan->d_dct, plan->nc_pad, lm, clipmax_dev + c0, nc, T, plan->cfg.n_mels,
                                      plan->cfg.n_mfcc, plan->cfg.top_db, mf, dl, clamp_in_place, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "mfcc_kernel launch");
  }
  return MMF_OK;
}

int mmf_sosfiltfilt(mmf_plan* plan, const void* x_dev, int32_t x_is_f32, int64_t rows, int64_t T,
                    int64_t x_row_stride, const double* sos_host, int32_t n_sections, double* y_dev,
                    int64_t y_row_stride, void* stream) {
  if (!plan || !x_dev || !sos_host || !y_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (rows < 1) return fail(MMF_ERR_INVALID, "rows must be positive");
  SosArgs a;
  int rc = fill_sos(sos_host, n_sections, &a);
  if (rc) return rc;
  if (T <= a.padlen) {
    char buf[128];
    std::snprintf(buf, sizeof(buf), "The length of the input vector x must be greater than padlen, which is %d.",
                  a.padlen);
    return fail(MMF_ERR_TOO_SHORT, buf);
  }
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  cudaError_t e = sosfiltfilt_any(x_dev, x_is_f32, rows, T, x_row_stride,
                                  (int)(rows > 0x7fffffffL ? 0x7fffffff : rows), 0, a, y_dev, y_row_stride,
                                  (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "sosfiltfilt_kernel launch");
  return MMF_OK;
}

int mmf_delta_norm(mmf_plan* plan, const double* x_dev, int64_t n_clips, int32_t rows_per_clip, int64_t T,
                   int32_t method, double* tot_dev, void* stream) {
  if (!plan || !x_dev || !tot_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (n_clips < 1 || rows_per_clip < 1 || T < 1) return fail(MMF_ERR_INVALID, "sizes must be positive");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  for (int64_t c0 = 0; c0 < n_clips; c0 += 65535) {
    const int64_t nc = std::min<int64_t>(65535, n_clips - c0);
    cudaError_t e = delta_norm_launch(x_dev + (size_t)c0 * rows_per_clip * T, nc, rows_per_clip, T, method,
                                      tot_dev + (size_t)c0 * T, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "delta_norm_kernel launch");
  }
  return MMF_OK;
}

// small host arrays are staged through the plan workspace tail (first 64 KB reserved)
static int stage_consts(mmf_plan* plan, const void* host, size_t bytes, size_t offset, void** dev, cudaStream_t st) {
  int rc = ensure_ws(plan, std::max<size_t>(plan->ws_bytes, 1 << 20));
  if (rc) return rc;
  if (offset + bytes > (64 << 10)) return fail(MMF_ERR_UNSUPPORTED, "coefficient arrays larger than 64 KB");
  *dev = (unsigned char*)plan->ws + offset;
  MMF_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, st));
  return MMF_OK;
}

int mmf_fir_filtfilt(mmf_plan* plan, const double* x_dev, int64_t rows, int64_t T, const double* b_host,
                     int32_t n_taps, double* y_dev, double* work_dev, void* stream) {
  if (!plan || !x_dev || !b_host || !y_dev || !work_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (n_taps < 1 || n_taps > 4096) return fail(MMF_ERR_UNSUPPORTED, "n_taps must be in [1, 4096]");
  if (rows < 1 || rows > 65535) return fail(MMF_ERR_UNSUPPORTED, "rows must be in [1, 65535]");
  if (T <= 3 * n_taps) {
    char buf[128];
    std::snprintf(buf, sizeof(buf), "The length of the input vector x must be greater than padlen, which is %d.",
                  3 * n_taps);
    return fail(MMF_ERR_TOO_SHORT, buf);
  }
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  void* b_dev = nullptr;
  int rc = stage_consts(plan, b_host, (size_t)n_taps * 8, 0, &b_dev, (cudaStream_t)stream);
  if (rc) return rc;
  cudaError_t e = fir_filtfilt_launch(x_dev, rows, T, (const double*)b_dev, n_taps, y_dev, work_dev,
                                      (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "fir kernels launch");
  return MMF_OK;
}

int mmf_stencil(mmf_plan* plan, const double* x_dev, int64_t rows, int64_t T, const double* coef_host, int32_t half,
                const double* edge_l_host, const double* edge_r_host, int32_t n_edge, int32_t n_edge_in,
                double* y_dev, void* stream) {
  if (!plan || !x_dev || !coef_host || !y_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (half < 0 || n_edge < half || n_edge_in < 0) return fail(MMF_ERR_INVALID, "need n_edge >= half >= 0");
  if (n_edge > 0 && (!edge_l_host || !edge_r_host)) return fail(MMF_ERR_INVALID, "edge matrices are NULL");
  if (T < n_edge_in || T < 2 * n_edge) return fail(MMF_ERR_TOO_SHORT, "input shorter than the stencil's edge window");
  if (rows < 1 || rows > 65535) return fail(MMF_ERR_UNSUPPORTED, "rows must be in [1, 65535]");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cb = (size_t)(2 * half + 1) * 8, eb = (size_t)n_edge * n_edge_in * 8;
  void *c_dev = nullptr, *l_dev = nullptr, *r_dev = nullptr;
  int rc = stage_consts(plan, coef_host, cb, 0, &c_dev, st);
  if (rc) return rc;
  if (eb) {
    if ((rc = stage_consts(plan, edge_l_host, eb, align_up(cb, 256), &l_dev, st))) return rc;
    if ((rc = stage_consts(plan, edge_r_host, eb, align_up(cb, 256) + align_up(eb, 256), &r_dev, st))) return rc;
  }
  cudaError_t e = stencil_launch(x_dev, rows, T, (const double*)c_dev, half, (const double*)l_dev,
                                 (const double*)r_dev, n_edge, n_edge_in, y_dev, st);
  if (e != cudaSuccess) return cuda_fail(e, "stencil_kernel launch");
  return MMF_OK;
}

int mmf_modspec(mmf_plan* plan, const float* mfcc_dev, int64_t n_clips, int32_t n_coef, int64_t T, int32_t win,
                int32_t hop, int32_t nfft, float* mag_dev, float* band_dev, const int32_t* band_lo_host,
                const int32_t* band_hi_host, int32_t n_bands, void* stream) {
  if (!plan || !mfcc_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (win < 1 || hop < 1 || nfft < win || nfft < 32 || nfft > 4096 || (nfft & (nfft - 1)))
    return fail(MMF_ERR_UNSUPPORTED, "need 1 <= win <= nfft, nfft a power of two in [32, 4096], hop >= 1");
  if (n_bands < 0 || n_bands > 16) return fail(MMF_ERR_UNSUPPORTED, "n_bands must be in [0, 16]");
  if (band_dev && n_bands > 0 && (!band_lo_host || !band_hi_host)) return fail(MMF_ERR_INVALID, "band edges are NULL");
  if (n_clips < 1 || n_coef < 1) return fail(MMF_ERR_INVALID, "n_clips and n_coef must be positive");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  if (band_dev)
    for (int b = 0; b < n_bands; ++b)
      if (band_lo_host[b] < 0 || band_hi_host[b] > nfft / 2 + 1 || band_lo_host[b] > band_hi_host[b])
        return fail(MMF_ERR_INVALID, "band bin range outside [0, nfft/2+1]");
  return run_modspec(plan, mfcc_dev, n_clips, n_coef, T, win, hop, nfft, mag_dev, band_dev, band_lo_host, band_hi_host,
                     n_bands, (cudaStream_t)stream);
}

int mmf_rms(mmf_plan* plan, const float* pcm_dev, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
            int32_t frame_length, int32_t hop_length, int32_t center, float* rms_dev, void* stream) {
  if (!plan || !pcm_dev || !rms_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (frame_length < 1 || hop_length < 1) return fail(MMF_ERR_INVALID, "frame_length and hop_length must be positive");
  const int pad = center ? frame_length / 2 : 0;
  const int64_t padded = n_samples + 2 * (int64_t)pad;
  if (padded < frame_length) return fail(MMF_ERR_TOO_SHORT, "Input is too short for frame_length");
  const int64_t T = 1 + (padded - frame_length) / hop_length;
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  for (int64_t c0 = 0; c0 < n_clips; c0 += 65535) {
    const int64_t nc = std::min<int64_t>(65535, n_clips - c0);
    cudaError_t e = rms_launch(pcm_dev + (size_t)c0 * clip_stride, nc, n_samples, clip_stride, frame_length,
                               hop_length, pad, T, rms_dev + (size_t)c0 * T, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "rms_kernel launch");
  }
  return MMF_OK;
}

}  // extern "C"

namespace mmf {

static size_t change_ws_bytes(const mmf_plan* p, int64_t n_clips, int64_t T, bool need_logmel, bool need_mfcc,
                              int rows) {
  size_t b = 64 << 10;
  if (need_logmel) b += align_up((size_t)n_clips * p->cfg.n_mels * T * 4, 256);
  b += align_up((size_t)n_clips * 4, 256);
  if (need_mfcc) b += align_up((size_t)n_clips * p->cfg.n_mfcc * T * 4, 256);
  b += align_up((size_t)n_clips * rows * T * 8, 256);
  b += align_up((size_t)n_clips * T * 8, 256);
  return b;
}

static int too_short(int padlen) {
  char buf[128];
  std::snprintf(buf, sizeof(buf), "The length of the input vector x must be greater than padlen, which is %d.", padlen);
  return fail(MMF_ERR_TOO_SHORT, buf);
}

// log-mel -> clamp -> MFCC (+delta) -> zero-phase Butterworth per coefficient ->
// derivative + norm -> output filter.  `cur` bumps through caller-provided workspace.
static int run_post(mmf_plan* p, unsigned char*& cur, float* logmel, const int* clipmax, int64_t n_clips, int64_t T,
                    const mmf_change_params* prm, double* tot, float* mfcc_out, float* delta_out, int clamp_in_place,
                    cudaStream_t st) {
  const mmf_config& c = p->cfg;
  const int first = prm->remove_first ? 1 : 0;
  const int rows = c.n_mfcc - first;
  if (rows < 1) return fail(MMF_ERR_INVALID, "removeFirst leaves no MFCC rows");
  SosArgs sa, so;
  int rc = fill_sos(prm->sos, prm->n_sections, &sa);
  if (rc) return rc;
  if (prm->out_kind == 0 && (rc = fill_sos(prm->out_sos, prm->out_n_sections, &so))) return rc;
  if (T <= sa.padlen) return too_short(sa.padlen);
  if (prm->out_kind == 0 && T <= so.padlen) return too_short(so.padlen);
  SosPar pa, po;
  size_t fsmem = 0;
  const bool out_iir = prm->out_kind == 0;
  const bool par_ok = !(c.flags & MMF_FLAG_UNFUSED_CHANGE) && sos_par_fill(sa, T, &pa) &&
                      (!out_iir || sos_par_fill(so, T, &po));
  // one kernel per clip from log-mel to the curve: clamp + DCT-II (K3) feed the float64 row buffers in
  // shared memory directly; MFCC / delta go to HBM only if the caller asked for them.  Measured on the
  // bench workload: 1 % faster than the separate MFCC kernel without a delta output, 1 % slower with
  // one when the modulation spectrum follows -- hence the default.
  // Batches: the output filter of the curve runs as its own launch (one warp per clip, all clips at once) instead of
  // on ONE warp of the per-clip CTA while its other 11 wait (27 % of that kernel's stall samples, ncu round 2); the
  // raw curve makes an 8 KB round trip per clip.  Same routine on the same buffer contents: bit-identical.
  const bool split_out = out_iir && par_ok && n_clips >= 32;
  const bool fold = (c.flags & MMF_FLAG_FOLD_MFCC) || delta_out == nullptr;
  if (par_ok && fold && !(c.flags & (MMF_FLAG_SEPARATE_MFCC | MMF_FLAG_MMA_DCT)) && n_clips <= 0x7fffffff &&
      change_fused_lm_supported(pa, out_iir ? &po : nullptr, c.n_mfcc, c.n_mels, first, rows, T, &fsmem)) {
    const FusedMfccArgs lm{p->d_dct, p->nc_pad, logmel, clipmax, c.n_mels, c.top_db, mfcc_out, delta_out,
                           clamp_in_place};
    double* raw = split_out ? (double*)carve(cur, (size_t)n_clips * T * 8) : nullptr;
    cudaError_t e = change_fused_launch(nullptr, n_clips, c.n_mfcc, first, rows, T, prm->diff_method, pa,
                                        out_iir ? po : pa, split_out ? 1 : prm->out_kind, split_out ? raw : tot, fsmem,
                                        &lm, st);
    if (e == cudaSuccess && split_out) e = sosfiltfilt_par_launch(raw, 0, n_clips, T, T, 1, T, po, tot, T, st);
    if (e == cudaSuccess) return MMF_OK;
    if (e != cudaErrorNotSupported) return cuda_fail(e, "change_fused_kernel (from log-mel) launch");
  }
  float* mfcc = mfcc_out ? mfcc_out : (float*)carve(cur, (size_t)n_clips * c.n_mfcc * T * 4);
  if ((rc = mmf_mfcc(p, logmel, clipmax, n_clips, T, mfcc, delta_out, clamp_in_place, st))) return rc;
  // fused per-clip kernel: row filters, derivative + norm and the output filter with the
  // float64 rows resident in shared memory (script/mfcc.py:393-425)
  if (par_ok && change_fused_supported(pa, out_iir ? &po : nullptr, rows, T, &fsmem)) {
    double* raw = split_out ? (double*)carve(cur, (size_t)n_clips * T * 8) : nullptr;
    cudaError_t e = change_fused_launch(mfcc, n_clips, c.n_mfcc, first, rows, T, prm->diff_method, pa,
                                        out_iir ? po : pa, split_out ? 1 : prm->out_kind, split_out ? raw : tot, fsmem,
                                        nullptr, st);
    if (e == cudaSuccess && split_out) e = sosfiltfilt_par_launch(raw, 0, n_clips, T, T, 1, T, po, tot, T, st);
    if (e == cudaSuccess) return MMF_OK;
    if (e != cudaErrorNotSupported) return cuda_fail(e, "change_fused_kernel launch");
  }
  double* filt = (double*)carve(cur, (size_t)n_clips * rows * T * 8);
  double* raw = (double*)carve(cur, (size_t)n_clips * T * 8);
  // Butterworth zero-phase low-pass of rows first..n_mfcc-1 of every clip (script/mfcc.py:398-402)
  cudaError_t e = sosfiltfilt_any(mfcc + (size_t)first * T, 1, n_clips * rows, T, T, rows, (long)c.n_mfcc * T, sa,
                                  filt, T, st);
  if (e != cudaSuccess) return cuda_fail(e, "sosfiltfilt (mfcc rows)");
  double* change = out_iir ? raw : tot;
  if ((rc = mmf_delta_norm(p, filt, n_clips, rows, T, prm->diff_method, change, st))) return rc;
  if (out_iir) {
    e = sosfiltfilt_any(raw, 0, n_clips, T, T, (int)std::min<int64_t>(n_clips, 0x7fffffff), 0, so, tot, T, st);
    if (e != cudaSuccess) return cuda_fail(e, "sosfiltfilt (total change)");
  }
  return MMF_OK;
}

// device-resident body; `base` is a workspace of change_ws_bytes()
static int run_change(mmf_plan* p, unsigned char* base, const float* pcm, int64_t n_clips, int64_t n_samples,
                      int64_t clip_stride, const mmf_change_params* prm, double* tot, float* logmel_out,
                      float* mfcc_out, float* delta_out, cudaStream_t st) {
  const mmf_config& c = p->cfg;
  const int64_t T = mmf_num_frames(n_samples, c.n_fft, c.hop_length);
  unsigned char* cur = base + (64 << 10);
  float* logmel = logmel_out ? logmel_out : (float*)carve(cur, (size_t)n_clips * c.n_mels * T * 4);
  int* clipmax = (int*)carve(cur, (size_t)n_clips * 4);
  int rc;
  if ((rc = run_stft(p, pcm, n_clips, n_samples, clip_stride, nullptr, logmel, clipmax, st))) return rc;
  return run_post(p, cur, logmel, clipmax, n_clips, T, prm, tot, mfcc_out, delta_out, logmel_out ? 1 : 0, st);
}

}  // namespace mmf

extern "C" {

int mmf_mfcc_change(mmf_plan* plan, const float* pcm_dev, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                    const mmf_change_params* prm, double* tot_dev, float* logmel_dev, float* mfcc_dev,
                    float* delta_dev, void* stream) {
  if (!plan || !pcm_dev || !prm || !tot_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (n_clips < 1 || n_samples < 1) return fail(MMF_ERR_INVALID, "n_clips and n_samples must be positive");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  const int64_t T = mmf_num_frames(n_samples, plan->cfg.n_fft, plan->cfg.hop_length);
  const int rows = plan->cfg.n_mfcc - (prm->remove_first ? 1 : 0);
  const size_t need = change_ws_bytes(plan, n_clips, T, logmel_dev == nullptr, mfcc_dev == nullptr, std::max(rows, 1));
  int rc = ensure_ws(plan, need);
  if (rc) return rc;
  return run_change(plan, (unsigned char*)plan->ws, pcm_dev, n_clips, n_samples, clip_stride, prm, tot_dev,
                    logmel_dev, mfcc_dev, delta_dev, (cudaStream_t)stream);
}

int mmf_change_from_logmel(mmf_plan* plan, float* logmel_dev, const int32_t* clipmax_dev, int64_t n_clips, int64_t T,
                           const mmf_change_params* prm, double* tot_dev, float* mfcc_dev, float* delta_dev,
                           int32_t clamp_in_place, void* stream) {
  if (!plan || !logmel_dev || !clipmax_dev || !prm || !tot_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (n_clips < 1 || T < 1) return fail(MMF_ERR_INVALID, "n_clips and T must be positive");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  const int rows = plan->cfg.n_mfcc - (prm->remove_first ? 1 : 0);
  const size_t need = change_ws_bytes(plan, n_clips, T, false, mfcc_dev == nullptr, std::max(rows, 1));
  int rc = ensure_ws(plan, need);
  if (rc) return rc;
  unsigned char* cur = (unsigned char*)plan->ws + (64 << 10);
  return run_post(plan, cur, logmel_dev, clipmax_dev, n_clips, T, prm, tot_dev, mfcc_dev, delta_dev, clamp_in_place,
                  (cudaStream_t)stream);
}

}  // extern "C"

// pcm_host: float32 samples, or (pcm16 != 0) int16 samples converted on the device
static int features_host_impl(mmf_plan* plan, const void* pcm_host_v, int pcm16, int64_t n_clips, int64_t n_samples,
                              int64_t clip_stride, const mmf_change_params* prm, const mmf_modspec_params* mod,
                              double* tot_host, float* mfcc_host, float* delta_host, float* mag_host,
                              float* band_host) {
  const float* pcm_host = (const float*)pcm_host_v;
  const int16_t* pcm_host16 = (const int16_t*)pcm_host_v;
  if (!plan || !pcm_host_v || !prm || !tot_host) return fail(MMF_ERR_INVALID, "NULL argument");
  if (n_clips < 1 || n_samples < 1) return fail(MMF_ERR_INVALID, "n_clips and n_samples must be positive");
  if ((mag_host || band_host) && !mod) return fail(MMF_ERR_INVALID, "modulation outputs requested without parameters");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  const mmf_config& c = plan->cfg;
  const int64_t T = mmf_num_frames(n_samples, c.n_fft, c.hop_length);
  if (T < 1) return fail(MMF_ERR_INVALID, "input too short for one frame");
  const int rows = std::max(1, c.n_mfcc - (prm->remove_first ? 1 : 0));
  const bool want_mod = mod && (mag_host || band_host) && mod->win > 0;
  int64_t n_win = 0;
  int nbins = 0;
  if (want_mod) {
    if (mod->hop < 1 || mod->nfft < mod->win || mod->nfft > 4096 || (mod->nfft & (mod->nfft - 1)) || mod->n_bands < 0 ||
        mod->n_bands > 16)
      return fail(MMF_ERR_UNSUPPORTED, "bad modulation-spectrum parameters");
    n_win = T >= mod->win ? 1 + (T - mod->win) / mod->hop : 0;
    nbins = mod->nfft / 2 + 1;
  }
  const bool need_mfcc_dev = mfcc_host || want_mod;
  // chunk the batch so that H2D of chunk i+1 overlaps compute of chunk i (two streams, two slots)
  const size_t clip_bytes = (size_t)n_samples * 4;
  // chunk size in bytes of float32 PCM on the device, from a sweep on B200 (bench e2e legs, 1024 x 10 s clips):
  // float32 host input 8 / 16 / 24 / 32 / 48 / 96 MB -> 0.68 / 0.81 / 0.82 / 0.82 / 0.81 / 0.80 M audio-s/s (small
  // chunks pay launch + copy set-up, large ones a longer unoverlapped head and tail); int16 host input moves half
  // the bytes per clip and keeps gaining up to 96 MB (0.94 / 1.16 / 1.25 / 1.34 / 1.37 / 1.44 M)
  const size_t chunk_bytes = pcm16 ? (96u << 20) : (32u << 20);
  int64_t chunk = std::max<int64_t>(1, (int64_t)(chunk_bytes / clip_bytes));
  chunk = std::min<int64_t>(chunk, n_clips);
  const size_t pcm_slot = align_up((size_t)chunk * n_samples * 4, 256);
  const size_t tot_slot = align_up((size_t)chunk * T * 8, 256);
  const size_t mfcc_slot = need_mfcc_dev ? align_up((size_t)chunk * c.n_mfcc * T * 4, 256) : 0;
  const size_t delta_slot = delta_host ? align_up((size_t)chunk * c.n_mfcc * T * 4, 256) : 0;
  const size_t mag_slot = (want_mod && mag_host) ? align_up((size_t)chunk * c.n_mfcc * n_win * nbins * 4, 256) : 0;
  const size_t band_slot = (want_mod && band_host) ? align_up((size_t)chunk * n_win * mod->n_bands * 4 + 4, 256) : 0;
  const size_t change_slot = change_ws_bytes(plan, chunk, T, true, !need_mfcc_dev, rows);
  const size_t raw_slot = pcm16 ? align_up((size_t)chunk * n_samples * 2, 256) : 0;  // int16 staging
  const size_t slot = pcm_slot + raw_slot + tot_slot + mfcc_slot + delta_slot + mag_slot + band_slot + change_slot;
  int rc = ensure_host_ws(plan, (64 << 10) + 2 * slot);
  if (rc) return rc;
  // the modulation tables are (re)built with a device-wide synchronisation: do it before any copy is queued
  if (want_mod && n_win > 0 &&
      (rc = ensure_mod_tables(plan, mod->win, mod->nfft, mod->band_lo, mod->band_hi, band_host ? mod->n_bands : 0)))
    return rc;
  // every exit below drains both streams: queued copies target caller-owned host buffers
  auto body = [&]() -> int {
  int64_t done = 0;
  for (int i = 0; done < n_clips; ++i, done += chunk) {
    const int s = i & 1;
    const int64_t nc = std::min<int64_t>(chunk, n_clips - done);
    cudaStream_t st = plan->streams[s];
    unsigned char* cur = (unsigned char*)plan->host_ws + (64 << 10) + (size_t)s * slot;
    float* d_pcm = (float*)cur;
    cur += pcm_slot;
    int16_t* d_raw = pcm16 ? (int16_t*)cur : nullptr;
    cur += raw_slot;
    double* d_tot = (double*)cur;
    cur += tot_slot;
    float* d_mfcc = need_mfcc_dev ? (float*)cur : nullptr;
    cur += mfcc_slot;
    float* d_delta = delta_host ? (float*)cur : nullptr;
    cur += delta_slot;
    float* d_mag = mag_slot ? (float*)cur : nullptr;
    cur += mag_slot;
    float* d_band = band_slot ? (float*)cur : nullptr;
    cur += band_slot;
    // stream order serialises reuse of slot s (chunk i-2 ran on the same stream)
    if (pcm16) {
      if (clip_stride == n_samples) {
        MMF_CUDA(cudaMemcpyAsync(d_raw, pcm_host16 + (size_t)done * clip_stride, (size_t)nc * n_samples * 2,
                                 cudaMemcpyHostToDevice, st));
      } else {
        MMF_CUDA(cudaMemcpy2DAsync(d_raw, (size_t)n_samples * 2, pcm_host16 + (size_t)done * clip_stride,
                                   (size_t)clip_stride * 2, (size_t)n_samples * 2, (size_t)nc, cudaMemcpyHostToDevice,
                                   st));
      }
      cudaError_t ce = pcm16_to_f32_launch(d_raw, (long)nc * n_samples, d_pcm, st);
      if (ce != cudaSuccess) return cuda_fail(ce, "pcm16_to_f32_kernel launch");
    } else if (clip_stride == n_samples) {
      MMF_CUDA(cudaMemcpyAsync(d_pcm, pcm_host + (size_t)done * clip_stride, (size_t)nc * n_samples * 4,
                               cudaMemcpyHostToDevice, st));
    } else {
      MMF_CUDA(cudaMemcpy2DAsync(d_pcm, (size_t)n_samples * 4, pcm_host + (size_t)done * clip_stride,
                                 (size_t)clip_stride * 4, (size_t)n_samples * 4, (size_t)nc, cudaMemcpyHostToDevice,
                                 st));
    }
    // run_change carves its own scratch after a 64 KB constants area
    rc = run_change(plan, cur - (64 << 10), d_pcm, nc, n_samples, n_samples, prm, d_tot, nullptr, d_mfcc, d_delta, st);
    if (rc) return rc;
    if (want_mod && n_win > 0) {
      if ((rc = run_modspec(plan, d_mfcc, nc, c.n_mfcc, T, mod->win, mod->hop, mod->nfft, d_mag, d_band, mod->band_lo,
                            mod->band_hi, mod->n_bands, st)))
        return rc;
    }
    MMF_CUDA(cudaMemcpyAsync(tot_host + (size_t)done * T, d_tot, (size_t)nc * T * 8, cudaMemcpyDeviceToHost, st));
    if (mfcc_host)
      MMF_CUDA(cudaMemcpyAsync(mfcc_host + (size_t)done * c.n_mfcc * T, d_mfcc, (size_t)nc * c.n_mfcc * T * 4,
                               cudaMemcpyDeviceToHost, st));
    if (delta_host)
      MMF_CUDA(cudaMemcpyAsync(delta_host + (size_t)done * c.n_mfcc * T, d_delta, (size_t)nc * c.n_mfcc * T * 4,
                               cudaMemcpyDeviceToHost, st));
    if (d_mag && n_win > 0)
      MMF_CUDA(cudaMemcpyAsync(mag_host + (size_t)done * c.n_mfcc * n_win * nbins, d_mag,
                               (size_t)nc * c.n_mfcc * n_win * nbins * 4, cudaMemcpyDeviceToHost, st));
    if (d_band && n_win > 0)
      MMF_CUDA(cudaMemcpyAsync(band_host + (size_t)done * n_win * mod->n_bands, d_band,
                               (size_t)nc * n_win * mod->n_bands * 4, cudaMemcpyDeviceToHost, st));
  }
  return MMF_OK;
  };
  rc = body();
  const cudaError_t e0 = cudaStreamSynchronize(plan->streams[0]);
  const cudaError_t e1 = cudaStreamSynchronize(plan->streams[1]);
  if (rc) return rc;
  if (e0 != cudaSuccess) return cuda_fail(e0, "host path, stream 0");
  if (e1 != cudaSuccess) return cuda_fail(e1, "host path, stream 1");
  return MMF_OK;
}

extern "C" {

int mmf_features_host(mmf_plan* plan, const float* pcm_host, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                      const mmf_change_params* prm, const mmf_modspec_params* mod, double* tot_host, float* mfcc_host,
                      float* delta_host, float* mag_host, float* band_host) {
  return features_host_impl(plan, pcm_host, 0, n_clips, n_samples, clip_stride, prm, mod, tot_host, mfcc_host,
                            delta_host, mag_host, band_host);
}

int mmf_features_host_pcm16(mmf_plan* plan, const int16_t* pcm16_host, int64_t n_clips, int64_t n_samples,
                            int64_t clip_stride, const mmf_change_params* prm, const mmf_modspec_params* mod,
                            double* tot_host, float* mfcc_host, float* delta_host, float* mag_host,
                            float* band_host) {
  return features_host_impl(plan, pcm16_host, 1, n_clips, n_samples, clip_stride, prm, mod, tot_host, mfcc_host,
                            delta_host, mag_host, band_host);
}

int mmf_hilbert_envelope(mmf_plan* plan, const float* x_dev, int64_t n_clips, int64_t n, int64_t x_stride,
                         float* amp_dev, int64_t amp_stride, void* stream) {
  if (!plan || !x_dev || !amp_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (n_clips < 1 || n < 1 || n > (1L << 26)) return fail(MMF_ERR_UNSUPPORTED, "need 1 <= n <= 2^26 samples");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  cudaError_t e;
  if (n > 4096 && hilbert_fft_supported(n)) {
    // O(n log n): Bluestein chirp transform over a power-of-two Stockham FFT, float64 (hilbert_fft.cu)
    e = hilbert_fft_launch(x_dev, n_clips, n, x_stride, amp_dev, amp_stride, (cudaStream_t)stream);
  } else {
    // short signals: the equivalent circular convolution with the discrete Hilbert kernel, evaluated directly
    e = hilbert_envelope_launch(x_dev, n_clips, n, x_stride, amp_dev, amp_stride, plan->sm_count, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "hilbert kernels launch");
  return MMF_OK;
}

int mmf_find_peaks(mmf_plan* plan, const double* x_dev, int64_t rows, int64_t T, int64_t row_stride, int32_t minima,
                   int32_t max_peaks, int32_t* idx_dev, int32_t* count_dev, void* stream) {
  if (!plan || !x_dev || !idx_dev || !count_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (rows < 1 || T < 1 || max_peaks < 1 || rows > (1L << 26)) return fail(MMF_ERR_INVALID, "bad sizes");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  cudaError_t e = find_peaks_launch(x_dev, rows, T, row_stride, minima, max_peaks, idx_dev, count_dev,
                                    (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "find_peaks_kernel launch");
  return MMF_OK;
}

int mmf_pcm16_to_f32(mmf_plan* plan, const int16_t* pcm16_dev, int64_t n, float* pcm_dev, void* stream) {
  if (!plan || !pcm16_dev || !pcm_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (n < 1) return fail(MMF_ERR_INVALID, "n must be positive");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  cudaError_t e = pcm16_to_f32_launch(pcm16_dev, n, pcm_dev, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "pcm16_to_f32_kernel launch");
  return MMF_OK;
}

int mmf_resample_poly(mmf_plan* plan, const float* x_dev, int64_t n_clips, int64_t n_in, int64_t x_stride,
                      const float* h_host, int32_t len_h, int32_t up, int32_t down, int64_t n_pre_remove, int64_t n_out,
                      float* y_dev, int64_t y_stride, void* stream) {
  if (!plan || !x_dev || !h_host || !y_dev) return fail(MMF_ERR_INVALID, "NULL argument");
  if (n_clips < 1 || n_clips > 65535 || n_in < 1 || n_out < 1) return fail(MMF_ERR_INVALID, "bad sizes");
  if (up < 1 || down < 1 || len_h < 1 || len_h > (1 << 22))
    return fail(MMF_ERR_UNSUPPORTED, "need up, down >= 1 and a filter of at most 2^22 taps");
  MMF_CUDA(cudaSetDevice(plan->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  float* h_dev = nullptr;  // stream-ordered scratch for the taps (pageable source: the copy is staged, so
                           // h_host may be reused as soon as this call returns)
  MMF_CUDA(cudaMallocAsync((void**)&h_dev, (size_t)len_h * 4, st));
  cudaError_t e = cudaMemcpyAsync(h_dev, h_host, (size_t)len_h * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = resample_poly_launch(x_dev, n_clips, n_in, x_stride, h_dev, len_h, up, down, n_pre_remove, n_out, y_stride,
                             y_dev, st);
  cudaFreeAsync(h_dev, st);
  if (e != cudaSuccess) return cuda_fail(e, "resample_poly_kernel launch");
  return MMF_OK;
}

int mmf_mfcc_change_host(mmf_plan* plan, const float* pcm_host, int64_t n_clips, int64_t n_samples,
                         int64_t clip_stride, const mmf_change_params* prm, double* tot_host, float* mfcc_host) {
  return mmf_features_host(plan, pcm_host, n_clips, n_samples, clip_stride, prm, nullptr, tot_host, mfcc_host, nullptr,
                           nullptr, nullptr);
}

}  // extern "C"
