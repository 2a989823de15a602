// CubePad backward, cube-tile kernel (fp32, small planes) — included by cubepad.cu.
//
// The training path (temporal_model/train_temporal.py:105-107,167-170) back-propagates through every
// CubePad of the ConvLSTM cell (model/clstm.py:57-64: [6N,2000,7,7], [6N,4000,7,7] x2 per step).
// gx[f,y,x] = gy at the pixel's interior copy + gy at every halo position that copied the pixel
// (the transpose of the forward gather). As in the forward cube-tile kernel the unit of work is all
// six faces of a run of channels of one cube — the set inside which CubePad is closed — but here the
// PADDED gradient is what is staged: six TMA bulk loads (one contiguous k*Ho*Wo chunk per face) per
// tile, `stages` tiles ahead, every gy element read from DRAM exactly once, no atomics.
//
// Tables in shared memory (cubepad_geom.h, cubepad_for_each_copy: the push table inverted). They depend only on
// the geometry and the channels per stage, so the launcher keeps one copy per (device, geometry) in device memory —
// built by a one-CTA launch of this kernel (tab_out) — and every CTA of a normal launch fetches them with one bulk
// copy (tab); without a cached copy (first sight of a geometry inside a stream capture) every CTA builds them itself.
// One word per input position (face, y, x) — the staged word of its interior copy, the number of halo
// positions that copied it and where their list starts — and a short list of 16-bit staged words for the halo
// copies (only the pixels within a pad width of a face edge have any: 12 % at H = 32). Halo copies are summed in
// the fixed (entry, u, v) order — the same order as cubepad_bwd_band_kernel, so both paths produce bit-identical,
// reproducible sums. A consumer thread owns fixed input positions and walks the channels: LDS (+ adds for edge
// pixels) -> STG, consecutive lanes on consecutive words of the gx plane. (Round 1 kept 14 B of tables per position
// — 89 KB at H = 32 — which left room for only two 55 KB stages; at 4 B per position a third stage fits.)
#pragma once
#include "common.cuh"
#include "cubepad_cube.cuh"
#include "cubepad_geom.h"
#include "tma.cuh"

namespace cp360 {

struct CubeBwdArgs {
  const float* gy;
  float* gx;
  uint32_t* work;       // {next chunk, finished CTAs} (zero at launch) or nullptr: chunks dealt round-robin
  int64_t n_chunks;     // N * cblocks
  int32_t C;
  int32_t kmax;         // channels per chunk (multiple of the bulk-copy channel quantum)
  int32_t cblocks;
  int32_t stages;
  int32_t stage_words;  // 6 * kmax * Ho * Wo
  int32_t lut_off;      // byte offset of the position table (uint32 [6*H*W]): word:16 | halo copies:4 | list start:12
  int32_t ent_off;      // byte offset of the halo lists (uint16 [6*(Ho*Wo - H*W)], staged word of channel 0)
  int32_t ring_off;     // byte offset of the staging ring
  FastDiv d_HW;         // e / (H*W)
  int32_t reg_pos;      // 0: never use the register-cached position walk (A/B)
  // position tables prebuilt in global memory (a pure function of the geometry and kmax; cached per device by the
  // launcher): {n_interior, n_border, 0, 0} | tab_words words = the shared-memory image [lut_off, ring_off) |
  // uint16 [6*H*W] positions, interior ones first. tab == nullptr: every CTA builds them itself.
  const uint32_t* tab;
  uint32_t* tab_out;    // != nullptr: this launch (one CTA) only builds the tables and writes them here
  int32_t tab_words;
};

constexpr int kBwdRegPos = 12;   // positions per consumer thread that the register-cached walk holds (6*32*32 / 512)
constexpr int kBwdRegTK = 2;     // ... used when a stage holds at most this many channels (faces of 25..32 px)

// TK > 0: kmax known at compile time (the channel walk of a full chunk is unrolled in batches of 8)
template <int TK>
__global__ void __launch_bounds__(kCubeMaxThreads, 1)
cubepad_bwd_cube_kernel(const CubeBwdArgs a, const __grid_constant__ CubePadGeom g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                 // [stages]
  uint64_t* empty = full + kCubeMaxStages;                                // [stages]
  int64_t* chunk_of = reinterpret_cast<int64_t*>(empty + kCubeMaxStages); // [stages] chunk id staged there, -1: end
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + a.lut_off);
  uint16_t* ent = reinterpret_cast<uint16_t*>(smem_raw + a.ent_off);
  const float* ring = reinterpret_cast<const float*>(smem_raw + a.ring_off);
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  const int n_in = 6 * HW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_cons = (int)blockDim.x - 32, n_cons_warps = n_cons >> 5;
  const int kmax = TK ? TK : a.kmax;
  const int fstride = kmax * HoWo;                                        // face stride in a stage
  const int pm = max(max(g.pl, g.pr), max(g.pt, g.pd));
  const bool have_tab = a.tab != nullptr, builder = a.tab_out != nullptr;
  uint64_t* tabbar = reinterpret_cast<uint64_t*>(chunk_of + kCubeMaxStages) + 1;   // after the two list counters

  pdl_trigger();
  // ---- producer state (thread 0). Dynamic dealing: the first chunk of a CTA is its own index (no round trip to the
  // counter in front of the first load), every further one gridDim.x + a ticket drawn one step ahead of its use
  int ps = 0;
  uint32_t pph = 0, ticket = blockIdx.x;
  int64_t pit = 0;
  bool pdone = false;
  auto produce = [&]() {                               // stage one chunk (or the end mark)
    if (pit >= a.stages) tma::mbar_wait(&empty[ps], pph ^ 1u);
    int64_t q = (int64_t)blockIdx.x + pit * gridDim.x;
    if (a.work) {
      q = (int64_t)ticket;
      if (q < a.n_chunks) ticket = gridDim.x + atomicAdd(a.work, 1u);
    }
    if (q >= a.n_chunks) {
      chunk_of[ps] = -1;
      tma::mbar_arrive(&full[ps]);                     // completes the phase: consumers see the end mark
      pdone = true;
      return;
    }
    chunk_of[ps] = q;
    const int64_t n = q / a.cblocks;
    const int c0 = (int)(q - n * a.cblocks) * kmax;
    const int kl = min(kmax, a.C - c0);
    const uint32_t bytes = (uint32_t)(kl * HoWo) * 4u;
    tma::mbar_expect_tx(&full[ps], 6u * bytes);
    float* dst = const_cast<float*>(ring) + (size_t)ps * a.stage_words;
    const float* src = a.gy + ((n * 6) * a.C + c0) * HoWo;
#pragma unroll
    for (int f = 0; f < 6; ++f)
      tma::bulk_load(dst + f * fstride, src + (int64_t)f * a.C * HoWo, bytes, &full[ps]);
    ++pit;
    if (++ps == a.stages) { ps = 0; pph ^= 1u; }
  };
  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      tma::mbar_init(&full[s], 1);
      // every consumer THREAD arrives on empty[s] (CP360_ARRIVE_PER_WARP: one lane per warp after __syncwarp —
      // equally correct under the PTX memory model and equally fast on B200 (profiles/README.md §racecheck),
      // but compute-sanitizer's racecheck only credits a barrier to the threads that arrive on it, so the
      // per-thread form is the one that verifies clean)
#ifdef CP360_ARRIVE_PER_WARP
      tma::mbar_init(&empty[s], n_cons_warps);
#else
      tma::mbar_init(&empty[s], n_cons);
#endif
    }
    tma::mbar_init(tabbar, 1);
    tma::fence_mbar_init();
    if (have_tab) {
      tma::mbar_expect_tx(tabbar, (uint32_t)a.tab_words * 4u);
      tma::bulk_load(smem_raw + a.lut_off, a.tab + 4, (uint32_t)a.tab_words * 4u, tabbar);
    }
    if (!builder && !have_tab) {
      // the first stages - 1 chunks stream in while the tables below are built; the last stage is their scratch
      pdl_wait();
      for (int i = 0; i + 1 < a.stages && !pdone; ++i) produce();
    }
  }
  __syncwarp();
  // ---- tables. Only pixels within a pad width of a face edge receive halo copies (at H = 32: 12 % of the positions,
  // two lanes of every warp of a raster walk): they are first compacted into a list (in the idle last stage), so the
  // plate walks below run on full warps.
  // The register-cached walk (2-channel stages) also takes the positions without copies as a list: its warps then
  // hold either only such positions (two loads, two stores each) or border positions.
  uint16_t* blist = reinterpret_cast<uint16_t*>(const_cast<float*>(ring) + (size_t)(a.stages - 1) * a.stage_words);
  uint16_t* ilist = blist + n_in;                      // 4 B * n_in of scratch <= one stage (6 * kmax * Ho * Wo floats)
  int* n_border = reinterpret_cast<int*>(chunk_of + kCubeMaxStages);
  int* n_interior = n_border + 1;
  const int ctid = tid - 32;
  const int64_t CHW = (int64_t)a.C * HW;
  const bool reg_cached = TK > 0 && TK <= kBwdRegTK && n_in <= kBwdRegPos * n_cons && a.reg_pos != 0;
  if (tid == 0) { *n_border = 0; *n_interior = 0; }
  __syncthreads();
  if (have_tab) {
    if (warp != 0) tma::mbar_wait(tabbar, 0);          // the producer warp goes straight to its loads
  } else {
  for (int e0 = warp * 32; e0 < n_in; e0 += (int)blockDim.x) {
    const int e = e0 + lane;
    bool border = false;
    if (e < n_in) {
      const int f = e / HW, r = e - f * HW;
      const int y = r / g.W, x = r - y * g.W;
      border = min(min(y, g.H - 1 - y), min(x, g.W - 1 - x)) < pm;
      lut[e] = 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, border);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(n_border, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (border) blist[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)e;
    if (reg_cached || builder) {                       // a warp's run of interior pixels stays contiguous: coalesced stores
      const unsigned mi = __ballot_sync(0xffffffffu, e < n_in && !border);
      int ibase = 0;
      if (lane == 0 && mi) ibase = atomicAdd(n_interior, __popc(mi));
      ibase = __shfl_sync(0xffffffffu, ibase, 0);
      if (e < n_in && !border) ilist[ibase + __popc(mi & ((1u << lane) - 1u))] = (uint16_t)e;
    }
  }
  __syncthreads();
  const int nb = *n_border;
  // Pass 1: halo copies per border position
  for (int i = tid; i < nb; i += blockDim.x) {
    const int e = blist[i];
    const int f = e / HW, r = e - f * HW;
    const int y = r / g.W, x = r - y * g.W;
    int cnt = 0;
    cubepad_for_each_copy(g, f, y, x, [&](int, int, int) { ++cnt; });
    lut[e] = (uint32_t)cnt;
  }
  __syncthreads();
  if (warp == 0) {                                     // exclusive scan over the border list, one contiguous segment per lane
    const int seg = (nb + 31) / 32;
    const int b = min(lane * seg, nb), e1 = min(b + seg, nb);
    int sum = 0;
    for (int i = b; i < e1; ++i) sum += (int)lut[blist[i]];
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    int run = incl - sum;
    for (int i = b; i < e1; ++i) {
      const int e = blist[i];
      const int c = (int)lut[e];
      lut[e] = (uint32_t)run | (uint32_t)c << 16;      // start | count (start < 4096, count < 16: checked on the host)
      run += c;
    }
  }
  __syncthreads();
  // Pass 2: the halo lists, then the final position words
  for (int i = tid; i < nb; i += blockDim.x) {
    const int e = blist[i];
    const int f = e / HW, r = e - f * HW;
    const int y = r / g.W, x = r - y * g.W;
    int o = (int)(lut[e] & 0xffffu);
    cubepad_for_each_copy(g, f, y, x, [&](int dface, int oy, int ox) {
      ent[o++] = (uint16_t)(dface * fstride + oy * g.Wo + ox);
    });
  }
  __syncthreads();                                     // the list starts above are read before the words change format
  for (int e = tid; e < n_in; e += blockDim.x) {
    const int f = e / HW, r = e - f * HW;
    const int y = r / g.W, x = r - y * g.W;
    const uint32_t sc = lut[e];
    const int start = (int)(sc & 0xffffu), cnt = (int)(sc >> 16);
    const uint32_t inner = (uint32_t)(f * fstride + (y + g.pt) * g.Wo + x + g.pl);
    lut[e] = inner | (uint32_t)cnt << 16 | (uint32_t)start << 20;
  }
  __syncthreads();
  }  // !have_tab
  if (builder) {                                       // hand the tables to the launcher's cache
    const int n_int = *n_interior, nbb = *n_border;
    if (tid == 0) { a.tab_out[0] = (uint32_t)n_int; a.tab_out[1] = (uint32_t)nbb; a.tab_out[2] = 0; a.tab_out[3] = 0; }
    const uint32_t* img = reinterpret_cast<const uint32_t*>(smem_raw + a.lut_off);
    for (int i = tid; i < a.tab_words; i += blockDim.x) a.tab_out[4 + i] = img[i];
    uint16_t* ol = reinterpret_cast<uint16_t*>(a.tab_out + 4 + a.tab_words);
    for (int i = tid; i < n_in; i += blockDim.x) ol[i] = i < n_int ? ilist[i] : blist[i - n_int];
    return;
  }
  // register-cached walk: slot i = ctid + k * n_cons of [interior list | border list]; bit k of no_halo: every lane of
  // this warp holds an interior position there
  uint32_t pos_lut[kBwdRegPos];
  int pos_dst[kBwdRegPos];                              // destination offset inside the cube's block (< 2^31: host check), -1: none
  unsigned no_halo = 0;
  if (TK > 0 && TK <= kBwdRegTK && warp != 0) {
    const int n_int = have_tab ? (int)__ldg(a.tab) : *n_interior;
    const uint16_t* olist = reinterpret_cast<const uint16_t*>(a.tab + 4 + a.tab_words);
#pragma unroll
    for (int k = 0; k < kBwdRegPos; ++k) {
      const int i = ctid + k * n_cons;
      pos_lut[k] = 0;
      pos_dst[k] = -1;
      if (reg_cached && ctid >= 0 && i < n_in) {
        const int e = have_tab ? (int)__ldg(olist + i) : (i < n_int ? ilist[i] : blist[i - n_int]);
        const int f = fdiv(e, a.d_HW);
        pos_lut[k] = lut[e];
        pos_dst[k] = f * (int)CHW + (e - f * HW);
      }
      if (__all_sync(0xffffffffu, reg_cached && ctid >= 0 && i < n_int)) no_halo |= 1u << k;
    }
  }
  if (!have_tab) __syncthreads();                      // every read of the scratch lists is done before stage stages-1 is loaded
  pdl_wait();

  if (warp == 0) {
    // ---------------- producer: draw chunks, stage them `stages` deep
    if (lane == 0) {
      while (!pdone) produce();
      if (a.work && atomicAdd(a.work + 1, 1u) == gridDim.x - 1) {   // last CTA: hand the pair back zeroed
        a.work[0] = 0;
        a.work[1] = 0;
        __threadfence();
      }
    }
    return;
  }
  // ---------------- consumers
  int s = 0;
  uint32_t ph = 0;
  while (true) {
    tma::mbar_wait(&full[s], ph);
    const int64_t q = chunk_of[s];
    if (q < 0) break;
    const int64_t n = q / a.cblocks;
    const int c0 = (int)(q - n * a.cblocks) * kmax;
    const int kl = min(kmax, a.C - c0);
    const float* in_s = ring + (size_t)s * a.stage_words;
    float* __restrict__ out = a.gx + ((n * 6) * a.C + c0) * HW;
    if (TK > 0 && TK <= kBwdRegTK && kl == TK && reg_cached) {
      // few channels per stage (32x32 faces: 2): the per-position words above cost more than the data moves,
      // so they live in registers across chunks (every thread owns the same positions in every chunk)
      constexpr int KR = TK > 0 && TK <= kBwdRegTK ? TK : 1;
      float* outj[KR];                                 // one materialised base pointer per channel plane: a store is then
#pragma unroll                                         // one IMAD.WIDE + STG instead of a re-derived 64-bit index
      for (int j = 0; j < KR; ++j) {
        outj[j] = out + j * HW;
        asm volatile("" : "+l"(outj[j]));
      }
#pragma unroll
      for (int k = 0; k < kBwdRegPos; ++k) {
        if (no_halo >> k & 1u) {                       // warp-uniform
          const float* sp0 = in_s + (pos_lut[k] & 0xffffu);   // no copies to add
          float acc[KR];
#pragma unroll
          for (int j = 0; j < KR; ++j) acc[j] = sp0[j * HoWo];
#pragma unroll
          for (int j = 0; j < KR; ++j) __stcs(outj[j] + pos_dst[k], acc[j]);
        } else if (pos_dst[k] >= 0) {
          const uint32_t l = pos_lut[k];
          const int n_halo = (int)((l >> 16) & 15u), o0 = (int)(l >> 20);
          const float* sp0 = in_s + (l & 0xffffu);
          const float* sp1 = in_s + ent[n_halo ? o0 : 0];           // first halo copy (most border positions have one)
          float acc[KR];
#pragma unroll
          for (int j = 0; j < KR; ++j) {
            acc[j] = sp0[j * HoWo];
            const float v = sp1[j * HoWo];
            if (n_halo) acc[j] += v;
          }
          if (n_halo > 1) {
#pragma unroll 1
            for (int o = o0 + 1; o < o0 + n_halo; ++o) {
              const float* sp = in_s + ent[o];
#pragma unroll
              for (int j = 0; j < KR; ++j) acc[j] += sp[j * HoWo];
            }
          }
#pragma unroll
          for (int j = 0; j < KR; ++j) __stcs(outj[j] + pos_dst[k], acc[j]);
        }
      }
    } else if (TK > 0 && kl == TK) {
      constexpr int KB = TK >= 8 ? 8 : (TK > 0 ? TK : 1);
#pragma unroll 1
      for (int e = ctid; e < n_in; e += n_cons) {
        const uint32_t l = lut[e];
        const int n_halo = (int)((l >> 16) & 15u), o0 = (int)(l >> 20);
        const int f = fdiv(e, a.d_HW);
        float* __restrict__ dp = out + (int64_t)f * CHW + (e - f * HW);
        const float* sp0 = in_s + (l & 0xffffu);
#pragma unroll
        for (int c8 = 0; c8 < TK; c8 += KB) {
          float acc[KB];
#pragma unroll
          for (int j = 0; j < KB; ++j) acc[j] = sp0[(c8 + j) * HoWo];
#pragma unroll 1
          for (int o = o0; o < o0 + n_halo; ++o) {
            const float* sp = in_s + ent[o] + c8 * HoWo;
#pragma unroll
            for (int j = 0; j < KB; ++j) acc[j] += sp[j * HoWo];
          }
#pragma unroll
          for (int j = 0; j < KB; ++j) __stcs(dp + (c8 + j) * HW, acc[j]);
        }
      }
    } else {
#pragma unroll 1
      for (int e = ctid; e < n_in; e += n_cons) {
        const uint32_t l = lut[e];
        const int n_halo = (int)((l >> 16) & 15u), o0 = (int)(l >> 20);
        const int f = fdiv(e, a.d_HW);
        float* __restrict__ dp = out + (int64_t)f * CHW + (e - f * HW);
        const int inner = (int)(l & 0xffffu);
#pragma unroll 1
        for (int cc = 0; cc < kl; ++cc) {
          float acc = in_s[inner + cc * HoWo];
          for (int o = o0; o < o0 + n_halo; ++o) acc += in_s[ent[o] + cc * HoWo];
          __stcs(dp + cc * HW, acc);
        }
      }
    }
#ifdef CP360_ARRIVE_PER_WARP
    __syncwarp();                                      // orders every lane's reads of the stage before lane 0's release
    if (lane == 0) tma::mbar_arrive(&empty[s]);
#else
    tma::mbar_arrive(&empty[s]);                       // release: this thread's reads of the stage are done
#endif
    if (++s == a.stages) { s = 0; ph ^= 1u; }
  }
}

}  // namespace cp360
