// kvc_sm100a.cu — kernels + C ABI (include/kvc.h) of the B200 KV-cache compression path.
//
// One launch compresses every layer of a call: grid = (B*H, n_layers); one CTA owns one
// (layer, batch, head) unit and runs
//   scan   (K1) stream the K rows of the selection region once, fp32 sum of squares, round the
//               norm to the cache dtype (torch.norm semantics), build the radix key in shared
//               memory and its 11-bit histogram on the fly;
//   select (K2) radix select over the on-chip keys, ties -> lowest index, ascending indices;
//   gather (K3) copy sink rows + selected rows + tail rows of K and V into the dense output.
// Several CTAs are resident per SM, so one unit's select overlaps its neighbours' HBM phases.
// Algorithmic HBM bytes per unit: e*D*(R + 4*C)  (SURVEY.md §8d).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: a no-op call unless a profiler injected itself

#include "kvc_device.cuh"
#include "kvc_fused_tma.cuh"
#include "kvc_slab.cuh"
#include "kvc_vote.cuh"

// The library is ONE source file, but its kernel instantiations compile serially: the build compiles it several
// times in parallel, each pass keeping one group of entry points (and only the kernels they launch), and links the
// objects.  KVC_PART is a bit mask: 1 = compress / select / norms entry points (+ the library-wide state) and the
// bf16 fused kernels, 2 = slab entry points and the bf16 in-place kernels, 4 = vote, 8 = slab append kernels,
// 16 = f16/f32 fused kernels, 32 = f16/f32 in-place kernels.
#ifndef KVC_PART
#define KVC_PART 63
#endif
#define KVC_HAS_CORE (KVC_PART & 1)
#define KVC_HAS_SLAB (KVC_PART & 2)
#define KVC_HAS_VOTE (KVC_PART & 4)
#define KVC_HAS_APPEND (KVC_PART & 8)
#define KVC_HAS_FUSED_OTHER (KVC_PART & 16)
#define KVC_HAS_SLAB_OTHER (KVC_PART & 32)

#define KVC_STR2(x) #x
#define KVC_STR(x) KVC_STR2(x)

namespace kvc {

constexpr int kSmemFixed = kHistBins * 4 + kMiscInts * 4;

// ------------------------------------------------------------------ standalone K1
// One warp step = 32/lpr rows; grid-stride over row groups of one [B,H,R] problem.
template <int DT>
__global__ void __launch_bounds__(256) kvc_norm_kernel(const char* __restrict__ k_in, int64_t sb, int64_t sh,
                                                       int64_t ss, int H, int row_lo, int R, int64_t total_rows,
                                                       int lpr, int cpl, typename Traits<DT>::Key* __restrict__ out) {
    using Tr = Traits<DT>;
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr;
    const int sub = lane % lpr, rw = lane / lpr;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp_global * rpw; g < total_rows; g += n_warps * rpw) {
        const int64_t idx = g + rw;  // flat (bh, r)
        const bool ok = rw < rpw && idx < total_rows;
        float acc = 0.f;
        if (ok) {
            const int64_t bh = idx / R;
            const int r = (int)(idx - bh * R);
            const int64_t b = bh / H, h = bh - b * H;
            const char* p = k_in + b * sb + h * sh + (int64_t)(row_lo + r) * ss + (int64_t)sub * 16;
            for (int c = 0; c < cpl; ++c) acc = Tr::sumsq(ldg128_stream(p + (int64_t)c * lpr * 16), acc);
        }
        acc = group_sum_pow2(acc, lpr);
        if (ok && sub == 0) out[idx] = (typename Tr::Key)Tr::to_raw(sqrtf(acc));
    }
}

// ------------------------------------------------------------------ standalone K2
// One CTA per score row: load scores -> ordered keys + histogram in shared memory -> select.
template <int DT, int NT>
__global__ void __launch_bounds__(NT) kvc_select_kernel(const typename Traits<DT>::Key* __restrict__ scores, int n,
                                                        int k, int largest, int32_t* __restrict__ idx_out) {
    using Tr = Traits<DT>;
    using Key = typename Tr::Key;
    constexpr int kShift0 = Tr::kKeyBits - kHistBits;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem);
    int32_t* misc = reinterpret_cast<int32_t*>(smem + kHistBins * 4);
    int32_t* sidx = reinterpret_cast<int32_t*>(smem + kSmemFixed);
    Key* keys = reinterpret_cast<Key*>(smem + kSmemFixed + (size_t)((k + 3) & ~3) * 4);
    const int tid = threadIdx.x;
    const Key* src = scores + (int64_t)blockIdx.x * n;
    for (int i = tid; i < kHistBins; i += NT) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const Key key = ordered_key<Key>((uint32_t)src[i], largest != 0);
        keys[i] = key;
        atomicAdd(&hist[(uint32_t)key >> kShift0], 1u);
    }
    __syncthreads();
    block_radix_select<Key, NT>(keys, n, k, hist, misc, sidx, 0);
    int32_t* dst = idx_out + (int64_t)blockIdx.x * k;
    for (int i = tid; i < k; i += NT) dst[i] = sidx[i];
}

// ------------------------------------------------------------------ host side
#if KVC_HAS_CORE
thread_local char g_last_error[256] = "";
std::atomic<int64_t> g_launches{0};
#else
extern thread_local char g_last_error[256];
extern std::atomic<int64_t> g_launches;
#endif
constexpr int kMaxSmemOptin = 227 * 1024;

static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return KVC_ERR_CUDA;
}

static int elem_bytes(int dtype) { return dtype == KVC_DTYPE_F32 ? 4 : 2; }
static int key_bytes(int dtype) { return dtype == KVC_DTYPE_F32 ? 4 : 2; }

using FusedFn = void (*)(const BatchDev);
constexpr int kNT = 512;  // threads of the stand-alone select kernel

// Entry points run on shape->device and leave the calling thread's current device as they found it: torch keeps its
// own idea of the current device, and a library that moved it would send later allocations and stream queries of a
// multi-GPU process (device_map-style placement) to the wrong GPU.
struct DeviceGuard {
    int prev = -1;
    int status = KVC_OK;
    explicit DeviceGuard(int device) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) {
            prev = -1;
            status = cuda_fail(e, "cudaGetDevice");
            return;
        }
        if (prev == device) {
            prev = -1;  // nothing to restore
            return;
        }
        e = cudaSetDevice(device);
        if (e != cudaSuccess) {
            prev = -1;
            status = cuda_fail(e, "cudaSetDevice");
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// One NVTX range per entry point (tracing): `ncu --nvtx --nvtx-include "kvc_compress_layers/"` or an
// nsys timeline then attribute every launch to the C-ABI call that made it.  Without a tool attached the push / pop
// are two calls through a null-checked function pointer.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

static int ensure_smem(const void* fn, size_t bytes) {
    // The attribute is sticky per function and context; raising it again is cheap.
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemOptin);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(max dynamic smem)");
    }
    return KVC_OK;
}

// Largest power of two dividing cpr, capped at 32 (generic-path lanes per row).
static int pow2_lanes(int cpr) {
    int l = 1;
    while (l < 32 && (cpr % (l * 2)) == 0) l *= 2;
    return l;
}

static size_t fused_smem_bytes(int dtype, int max_region, int idx_cap) {
    const size_t keys = ((size_t)max_region * key_bytes(dtype) + 15) & ~(size_t)15;
    return (size_t)kSmemFixed + (size_t)idx_cap * 4 + keys;
}


// ------------------------------------------------------------------ TMA form: variant + smem plan
constexpr int kSmemPerSM = 228 * 1024;  // per-SM shared memory; every resident CTA reserves 1 KB of it
constexpr int kSMs = 148;               // B200
constexpr int kTmaHead = kMiscInts * 4 + 32 * 8;  // misc scalars + one mbarrier per warp

struct TmaPlan {
    bool ok = false;
    int nt = 0, ctas = 0, nsw = 0, upc = 1;
    int slide_slots = 0;  // in-place plans: slots the slide can use (planned slots + the dead key array)
    int off_hist = 0, off_idx = 0, off_keys = 0, off_stage = 0;
    size_t smem = 0;
};

// Launch-shape overrides and stage isolation exist only in a lab build (-DKVC_LAB: scripts/build_lab.sh); the product
// library reads no environment variable, so nothing outside the call's arguments can change what it computes.
#ifdef KVC_LAB
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
// KVC_VOTE_DEBUG switches stages of the vote kernel off for stage-isolation timing (1 = no math, 2 = no math, no
// MMA, 3 = no tail box, 4 = one K step): the votes it produces are NOT valid.
static int vote_debug_mode() { return env_int("KVC_VOTE_DEBUG", 0); }
#else
static inline int env_int(const char*, int dflt) { return dflt; }
static inline int vote_debug_mode() { return 0; }
#endif

// Choose threads per CTA, resident CTAs per SM and staging warps so that (a) the on-chip key
// buffer fits, (b) as close to 128 KB of rows as possible are in flight per SM, (c) as many CTAs as possible are
// resident so one unit's select phase hides under its neighbours' HBM phases.
// `units` / `unit_bytes` (optional): CTAs of the launch and the bytes one of them moves.  A launch of a few large
// units (c1: 960 units of 2.75 MB on 148 SMs) ends in a partial wave whose CTAs each run alone on their SM; with 3
// resident CTAs per SM that wave is a third of the whole launch, with one 512-thread CTA per SM a seventh
// (measured: 446 -> 402 us, 0.90 -> 1.00 of the copy peak, profiles/r02_plan_sweep_few_units.json).  Small units
// (decode steady state) and launches of many waves keep the residency-first choice.
// `in_place` (kvc_slab_compress with the keys on chip): the kept-index list is only written by the select's last pass,
// after the histogram's last reader, so it ALIASES the histogram (k_sel <= 2048); and the slide runs after the select,
// when the key array is dead and serves as staging, so the plan may hold NO slot of its own.  At 32K bf16 rows that is
// 72.7 KB instead of 84.6 KB per CTA: three resident CTAs per SM instead of two (one unit's select and slide rounds
// hide under its neighbours').
static TmaPlan plan_tma(int dtype, int cpr, int max_region, int idx_cap, bool any_select, bool light_traffic = false,
                        int64_t units = 0, int64_t unit_bytes = 0, bool in_place = false) {
    TmaPlan best;
    const int stage = 32 * cpr * 16;
    int end = kTmaHead;
    TmaPlan base;
    base.off_hist = base.off_idx = base.off_keys = kTmaHead;
    if (any_select) {
        base.off_hist = kTmaHead;
        base.off_idx = base.off_hist + kHistBins * 4;
        base.off_keys = base.off_idx + ((idx_cap * 4 + 15) & ~15);
        if (in_place && idx_cap <= kHistBins) {
            base.off_idx = base.off_hist;
            base.off_keys = base.off_hist + kHistBins * 4;
        }
        end = base.off_keys + (int)(((size_t)max_region * key_bytes(dtype) + 15) & ~(size_t)15);
    }
    base.off_stage = (end + 127) & ~127;
    const int dead_keys = in_place ? base.off_stage - ((base.off_keys + 127) & ~127) : 0;  // bytes the slide inherits
    const int force_nt = env_int("KVC_TMA_NT", 0), force_ctas = env_int("KVC_TMA_CTAS", 0),
              force_nsw = env_int("KVC_TMA_NSW", 0);
    static const int cand[][2] = {{256, 3}, {256, 2}, {512, 1}, {256, 1}};
    // a launch of a few large units: fewer than 8 waves at the highest residency that fits (then EVERY candidate is
    // ranked by estimated launch length, below)
    bool few_large_units = false;
    if (!light_traffic && units > 0 && unit_bytes >= (1 << 20) && !force_nt && !force_ctas) {
        for (const auto& c : cand) {
            int limit = kSmemPerSM / c[1] - 1024;
            if (limit > kMaxSmemOptin) limit = kMaxSmemOptin;
            if (limit - base.off_stage >= stage) {
                few_large_units = units / ((int64_t)c[1] * kSMs) < 8;
                break;  // candidates are listed by falling residency
            }
        }
    }
    long best_score = -1;
    for (const auto& c : cand) {
        const int nt = c[0], ctas = c[1];
        if (force_nt && nt != force_nt) continue;
        if (force_ctas && ctas != force_ctas) continue;
        int limit = kSmemPerSM / ctas - 1024;
        if (limit > kMaxSmemOptin) limit = kMaxSmemOptin;
        const int avail = limit - base.off_stage;
        if (avail < (in_place ? 0 : stage)) continue;
        int nsw = avail / stage;
        if (nsw > nt / 32) nsw = nt / 32;
        int slide = (dead_keys + nsw * stage) / stage;  // what the slab kernel computes for its rounds
        if (slide > nt / 32) slide = nt / 32;
        if (in_place && slide < 1) continue;
        if (force_nsw && force_nsw < nsw) nsw = force_nsw;
        const long inflight = (long)ctas * nsw * stage;
        // saturating score: bytes in flight up to 128 KB matter most (measured: c5 +8% from 64 -> 128 KB), then residency
        long score = (inflight < 131072 ? inflight : 131072) * 8 + ctas * 4096 + (inflight >> 6);
        // in-place compaction moves few bytes (scores come from stored norms): residency first, 2+ slots are enough
        if (light_traffic) score = ((in_place ? slide : nsw) >= 2 ? 1 : 0) * (1L << 30) + ctas * (1L << 20) + (in_place ? slide : nsw);
        if (few_large_units) {
            // rank by the estimated length of the launch: full waves plus a last partial wave that costs at least
            // ~0.35 of a full one (a lone CTA is latency-bound), over the rate this shape sustains — bytes in flight
            // (64 KB per SM measured 8 % behind 128 KB) and residency (3 CTAs hide a unit's select under its
            // neighbours' HBM phases: ~5 % on many-wave launches)
            const long slots = (long)ctas * kSMs;
            const long full = (long)(units / slots), rem = (long)(units - full * slots);
            const double est = (double)full * slots + (rem > 0 ? std::max((double)rem, 0.35 * slots) : 0.0);
            const double fill = (double)(inflight < 131072 ? inflight : 131072) / 131072.0;
            const double rate = (0.84 + 0.16 * fill) * (ctas >= 3 ? 1.05 : (ctas == 2 ? 1.03 : (nt == 512 ? 1.0 : 0.95)));
            score = (long)(1e12 / (est / rate + 1.0));
        }
        if (score > best_score) {
            best_score = score;
            best = base;
            best.ok = true;
            best.nt = nt;
            best.ctas = ctas;
            best.nsw = nsw;
            best.slide_slots = in_place ? slide : nsw;
            best.smem = (size_t)base.off_stage + (size_t)nsw * stage;
        }
    }
    best.upc = env_int("KVC_TMA_UPC", 1);
    if (best.upc < 1) best.upc = 1;
    return best;
}

// Keys + kept indices of one (layer, batch, head) unit when they live in the device workspace.
struct WsLayout {
    int64_t keys = 0, unit = 0;
};
static WsLayout ws_layout(int dtype, int max_region, int idx_cap) {
    WsLayout w;
    w.keys = ((int64_t)max_region * key_bytes(dtype) + 15) & ~(int64_t)15;
    w.unit = w.keys + (((int64_t)idx_cap * 4 + 15) & ~(int64_t)15);
    return w;
}
// The on-chip plan is good enough when it exists and keeps >= 48 KB of rows in flight per SM (scan kernels)
// or has at least two slots (in-place compaction, whose traffic is light).
static bool onchip_plan_ok(const TmaPlan& tp, int cpr, bool light_traffic) {
    if (!tp.ok) return false;
    if (light_traffic) return tp.slide_slots >= 2;
    return (long)tp.ctas * tp.nsw * 32 * cpr * 16 >= 48 * 1024;
}

#if KVC_HAS_CORE || KVC_HAS_FUSED_OTHER
template <int DT, int NT, int MINB>
static FusedFn pick_tma_cpr(int cpr) {
    switch (cpr) {
        case 8: return kvc_fused_tma_kernel<DT, 8, NT, MINB>;
        case 10: return kvc_fused_tma_kernel<DT, 10, NT, MINB>;
        case 12: return kvc_fused_tma_kernel<DT, 12, NT, MINB>;
        case 16: return kvc_fused_tma_kernel<DT, 16, NT, MINB>;
        case 20: return kvc_fused_tma_kernel<DT, 20, NT, MINB>;
        case 32: return kvc_fused_tma_kernel<DT, 32, NT, MINB>;
        default: return kvc_fused_tma_kernel<DT, 0, NT, MINB>;  // any other row width: run-time loops
    }
}
template <int DT>
static FusedFn pick_tma_nt(int cpr, int nt) {
    return nt == 512 ? pick_tma_cpr<DT, 512, 1>(cpr) : pick_tma_cpr<DT, 256, 3>(cpr);
}
#endif
FusedFn pick_tma_bf16(int cpr, int nt);
FusedFn pick_tma_other(int dtype, int cpr, int nt);
#if KVC_HAS_CORE
FusedFn pick_tma_bf16(int cpr, int nt) { return pick_tma_nt<KVC_DTYPE_BF16>(cpr, nt); }
static FusedFn pick_tma(int dtype, int cpr, int nt) {
    return dtype == KVC_DTYPE_BF16 ? pick_tma_bf16(cpr, nt) : pick_tma_other(dtype, cpr, nt);
}
#endif
#if KVC_HAS_FUSED_OTHER
FusedFn pick_tma_other(int dtype, int cpr, int nt) {
    return dtype == KVC_DTYPE_F32 ? pick_tma_nt<KVC_DTYPE_F32>(cpr, nt) : pick_tma_nt<KVC_DTYPE_F16>(cpr, nt);
}
#endif

// Rows of 16 B .. 2 KB: 128/160/192/256/320/512-byte rows have kernels with compile-time widths, the rest run the
// generic-width instantiation of the same kernels.
static bool row_width_ok(int cpr) { return cpr >= 1 && cpr <= 128; }

static int ensure_tma_attrs(const void* fn, int device) {
    // Function attributes are sticky per (function, device): set them once, not on every decode step.
    static std::mutex mu;
    static std::vector<std::pair<const void*, int>> done;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (const auto& d : done)
            if (d.first == fn && d.second == device) return KVC_OK;
    }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemOptin);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(max dynamic smem)");
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(carveout)");
    std::lock_guard<std::mutex> lock(mu);
    done.emplace_back(fn, device);
    return KVC_OK;
}


// ------------------------------------------------------------------ slab kernels: variant tables
using SlabFn = void (*)(const SlabBatchDev);
using AppendFn = void (*)(const AppendBatchDev);
using AppendOneFn = void (*)(const AppendOneDev);
SlabFn pick_slab_bf16(int cpr, int nt);
SlabFn pick_slab_other(int dtype, int cpr, int nt);
AppendFn pick_append(int dtype, int cpr);          // thread-per-row form, all layers
AppendOneFn pick_append_one(int dtype, int cpr);   // thread-per-row form, one layer (small parameter block)
AppendFn pick_append_tma(int dtype, int cpr);      // bulk-copy form (prefill-sized appends)

#if KVC_HAS_SLAB || KVC_HAS_SLAB_OTHER
template <int DT, int NT, int MINB>
static SlabFn pick_slab_cpr(int cpr) {
    switch (cpr) {
        case 8: return kvc_slab_compress_kernel<DT, 8, NT, MINB>;
        case 10: return kvc_slab_compress_kernel<DT, 10, NT, MINB>;
        case 12: return kvc_slab_compress_kernel<DT, 12, NT, MINB>;
        case 16: return kvc_slab_compress_kernel<DT, 16, NT, MINB>;
        case 20: return kvc_slab_compress_kernel<DT, 20, NT, MINB>;
        case 32: return kvc_slab_compress_kernel<DT, 32, NT, MINB>;
        default: return kvc_slab_compress_kernel<DT, 0, NT, MINB>;
    }
}
template <int DT>
static SlabFn pick_slab_nt(int cpr, int nt) {
    return nt == 512 ? pick_slab_cpr<DT, 512, 1>(cpr) : pick_slab_cpr<DT, 256, 3>(cpr);
}
#endif
#if KVC_HAS_SLAB
SlabFn pick_slab_bf16(int cpr, int nt) { return pick_slab_nt<KVC_DTYPE_BF16>(cpr, nt); }
static SlabFn pick_slab(int dtype, int cpr, int nt) {
    return dtype == KVC_DTYPE_BF16 ? pick_slab_bf16(cpr, nt) : pick_slab_other(dtype, cpr, nt);
}
#endif
#if KVC_HAS_SLAB_OTHER
SlabFn pick_slab_other(int dtype, int cpr, int nt) {
    return dtype == KVC_DTYPE_F32 ? pick_slab_nt<KVC_DTYPE_F32>(cpr, nt) : pick_slab_nt<KVC_DTYPE_F16>(cpr, nt);
}
#endif

#if KVC_HAS_APPEND
template <int DT, typename Params>
static void (*pick_append_row_cpr(int cpr))(const Params) {
    switch (cpr) {
        case 8: return kvc_slab_append_kernel<DT, 8, Params>;
        case 10: return kvc_slab_append_kernel<DT, 10, Params>;
        case 12: return kvc_slab_append_kernel<DT, 12, Params>;
        case 16: return kvc_slab_append_kernel<DT, 16, Params>;
        case 20: return kvc_slab_append_kernel<DT, 20, Params>;
        case 32: return kvc_slab_append_kernel<DT, 32, Params>;
        default: return kvc_slab_append_kernel<DT, 0, Params>;
    }
}
template <int DT>
static AppendFn pick_append_tma_cpr(int cpr) {
    switch (cpr) {
        case 8: return kvc_slab_append_tma_kernel<DT, 8>;
        case 10: return kvc_slab_append_tma_kernel<DT, 10>;
        case 12: return kvc_slab_append_tma_kernel<DT, 12>;
        case 16: return kvc_slab_append_tma_kernel<DT, 16>;
        case 20: return kvc_slab_append_tma_kernel<DT, 20>;
        case 32: return kvc_slab_append_tma_kernel<DT, 32>;
        default: return kvc_slab_append_tma_kernel<DT, 0>;
    }
}
AppendOneFn pick_append_one(int dtype, int cpr) {
    switch (dtype) {
        case KVC_DTYPE_F32: return pick_append_row_cpr<KVC_DTYPE_F32, AppendOneDev>(cpr);
        case KVC_DTYPE_F16: return pick_append_row_cpr<KVC_DTYPE_F16, AppendOneDev>(cpr);
        default: return pick_append_row_cpr<KVC_DTYPE_BF16, AppendOneDev>(cpr);
    }
}
AppendFn pick_append(int dtype, int cpr) {
    switch (dtype) {
        case KVC_DTYPE_F32: return pick_append_row_cpr<KVC_DTYPE_F32, AppendBatchDev>(cpr);
        case KVC_DTYPE_F16: return pick_append_row_cpr<KVC_DTYPE_F16, AppendBatchDev>(cpr);
        default: return pick_append_row_cpr<KVC_DTYPE_BF16, AppendBatchDev>(cpr);
    }
}
AppendFn pick_append_tma(int dtype, int cpr) {
    switch (dtype) {
        case KVC_DTYPE_F32: return pick_append_tma_cpr<KVC_DTYPE_F32>(cpr);
        case KVC_DTYPE_F16: return pick_append_tma_cpr<KVC_DTYPE_F16>(cpr);
        default: return pick_append_tma_cpr<KVC_DTYPE_BF16>(cpr);
    }
}
#endif  // KVC_HAS_APPEND

static int check_shape(const kvc_shape* shape, int* cpr_out) {
    if (!shape) return KVC_ERR_INVALID_ARG;
    const int B = shape->batch, H = shape->heads, D = shape->head_dim, dt = shape->dtype;
    if (B <= 0 || H <= 0 || D <= 0) return KVC_ERR_INVALID_ARG;
    if (dt != KVC_DTYPE_F32 && dt != KVC_DTYPE_F16 && dt != KVC_DTYPE_BF16) return KVC_ERR_UNSUPPORTED;
    const int e = elem_bytes(dt);
    if (((int64_t)D * e) % 16 != 0) return KVC_ERR_UNSUPPORTED;
    if ((int64_t)B * H > 0x7fffffffLL) return KVC_ERR_INVALID_ARG;
    *cpr_out = D * e / 16;
    return KVC_OK;
}

}  // namespace kvc

using namespace kvc;

extern "C" {

#if KVC_HAS_CORE
int kvc_abi_version(void) { return KVC_ABI_VERSION; }

const char* kvc_build_info(void) {
    return "libkvc_sm100a: sm_100a, nvcc " KVC_STR(__CUDACC_VER_MAJOR__) "." KVC_STR(__CUDACC_VER_MINOR__);
}

const char* kvc_status_string(int status) {
    switch (status) {
        case KVC_OK: return "ok";
        case KVC_ERR_INVALID_ARG: return "invalid argument";
        case KVC_ERR_UNSUPPORTED: return "unsupported dtype / head_dim / alignment";
        case KVC_ERR_TOO_LARGE: return "selection region too large for the on-chip score buffer";
        case KVC_ERR_CUDA: return "CUDA error";
        default: return "unknown status";
    }
}

const char* kvc_last_cuda_error(void) { return g_last_error; }

int64_t kvc_launch_count(void) { return g_launches.load(); }

int32_t kvc_max_region_rows(int32_t dtype, int32_t k_sel) {
    if (dtype < 0 || dtype > 2 || k_sel < 0) return 0;
    const int64_t idx = ((int64_t)k_sel + 3) & ~(int64_t)3;
    const int64_t left = (int64_t)kMaxSmemOptin - kSmemFixed - idx * 4;
    if (left <= 0) return 0;
    const int64_t rows = (left & ~(int64_t)15) / key_bytes(dtype);
    return (int32_t)(rows > 0x7fffffff ? 0x7fffffff : rows);
}

int kvc_compress_layers(const kvc_shape* shape, int32_t n_layers, const kvc_layer_plan* plans,
                        const kvc_layer_io* io, void* stream) {
    return kvc_compress_layers_ws(shape, n_layers, plans, io, nullptr, 0, stream);
}

// Launch-chunk statistics shared by kvc_workspace_bytes and the launchers.
static void chunk_stats(const kvc_layer_plan* plans, int nl, int* n_active, int* max_region, int* max_ksel,
                        bool* any_select, int64_t* rows_moved = nullptr) {
    *n_active = *max_region = *max_ksel = 0;
    *any_select = false;
    if (rows_moved) *rows_moved = 0;
    for (int l = 0; l < nl; ++l) {
        const kvc_layer_plan& p = plans[l];
        if (p.sink + p.k_sel + p.tail == 0) continue;
        ++*n_active;
        // rows one unit of this layer moves if it scans K: the region once + kept rows of K and V read and written
        if (rows_moved) *rows_moved += (p.k_sel > 0 ? p.sel_hi - p.sel_lo : 0) + 4LL * (p.sink + p.k_sel + p.tail);
        if (p.k_sel > 0) {
            *any_select = true;
            if (p.k_sel > *max_ksel) *max_ksel = p.k_sel;
            if (p.score != KVC_SCORE_GIVEN_INDEX && p.sel_hi - p.sel_lo > *max_region) *max_region = p.sel_hi - p.sel_lo;
        }
    }
}

int64_t kvc_workspace_bytes(const kvc_shape* shape, int32_t n_layers, const kvc_layer_plan* plans) {
    int cpr = 0;
    if (check_shape(shape, &cpr) != KVC_OK || n_layers <= 0 || !plans || !row_width_ok(cpr)) return 0;
    int64_t need = 0;
    for (int l0 = 0; l0 < n_layers; l0 += KVC_MAX_LAYERS_PER_LAUNCH) {
        const int nl = (n_layers - l0) < KVC_MAX_LAYERS_PER_LAUNCH ? (n_layers - l0) : KVC_MAX_LAYERS_PER_LAUNCH;
        int n_active, max_region, max_ksel;
        bool any_select;
        int64_t rows_moved;
        chunk_stats(plans + l0, nl, &n_active, &max_region, &max_ksel, &any_select, &rows_moved);
        if (!any_select) continue;
        const int idx_cap = (max_ksel + 3) & ~3;
        const TmaPlan tp = plan_tma(shape->dtype, cpr, max_region, idx_cap, true, false,
                                    (int64_t)shape->batch * shape->heads * n_active, rows_moved / n_active * cpr * 16);
        if (onchip_plan_ok(tp, cpr, false)) continue;
        const int64_t bytes = ws_layout(shape->dtype, max_region, idx_cap).unit * shape->batch * shape->heads * n_active;
        if (bytes > need) need = bytes;
    }
    return need;
}

int kvc_launch_shape(const kvc_shape* shape, int32_t n_layers, const kvc_layer_plan* plans, int32_t out[4]) {
    int cpr = 0;
    int st = check_shape(shape, &cpr);
    if (st != KVC_OK) return st;
    if (n_layers <= 0 || !plans || !out) return KVC_ERR_INVALID_ARG;
    if (!row_width_ok(cpr)) return KVC_ERR_UNSUPPORTED;
    const int nl = n_layers < KVC_MAX_LAYERS_PER_LAUNCH ? n_layers : KVC_MAX_LAYERS_PER_LAUNCH;
    int n_active, max_region, max_ksel;
    bool any_select;
    int64_t rows_moved;
    chunk_stats(plans, nl, &n_active, &max_region, &max_ksel, &any_select, &rows_moved);
    if (n_active == 0) return KVC_ERR_INVALID_ARG;
    const int idx_cap = (max_ksel + 3) & ~3;
    TmaPlan tp = plan_tma(shape->dtype, cpr, max_region, idx_cap, any_select, false,
                          (int64_t)shape->batch * shape->heads * n_active, rows_moved / n_active * cpr * 16);
    const bool ws = any_select && !onchip_plan_ok(tp, cpr, false);
    if (ws) tp = plan_tma(shape->dtype, cpr, 0, 0, true, false);
    if (!tp.ok) return KVC_ERR_TOO_LARGE;
    out[0] = tp.nt;
    out[1] = tp.ctas;
    out[2] = tp.nsw;
    out[3] = ws ? 1 : 0;
    return KVC_OK;
}

int kvc_compress_layers_ws(const kvc_shape* shape, int32_t n_layers, const kvc_layer_plan* plans,
                           const kvc_layer_io* io, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!shape || n_layers < 0 || (n_layers > 0 && (!plans || !io))) return KVC_ERR_INVALID_ARG;
    if (n_layers == 0) return KVC_OK;
    int cpr = 0;
    int st = check_shape(shape, &cpr);
    if (st != KVC_OK) return st;
    if (!row_width_ok(cpr)) return KVC_ERR_UNSUPPORTED;
    const int B = shape->batch, H = shape->heads, dt = shape->dtype;
    const int e = elem_bytes(dt);

    // validate every layer before anything is enqueued
    for (int l = 0; l < n_layers; ++l) {
        const kvc_layer_plan& p = plans[l];
        const kvc_layer_io& x = io[l];
        const int64_t C = (int64_t)p.sink + p.k_sel + p.tail;
        if (p.seq_len < 0 || p.sink < 0 || p.k_sel < 0 || p.tail < 0) return KVC_ERR_INVALID_ARG;
        if (p.sink > p.seq_len || p.tail > p.seq_len) return KVC_ERR_INVALID_ARG;
        if (p.k_sel > 0) {
            if (p.sel_lo < 0 || p.sel_hi > p.seq_len || p.sel_lo > p.sel_hi) return KVC_ERR_INVALID_ARG;
            if (p.k_sel > p.sel_hi - p.sel_lo) return KVC_ERR_INVALID_ARG;
            if (p.score <= KVC_SCORE_NONE || p.score > KVC_SCORE_GIVEN_SCORE) return KVC_ERR_INVALID_ARG;
            if (p.score == KVC_SCORE_GIVEN_INDEX && !x.idx_in) return KVC_ERR_INVALID_ARG;
            if (p.score == KVC_SCORE_GIVEN_SCORE && !x.score_in) return KVC_ERR_INVALID_ARG;
            if ((p.score == KVC_SCORE_GIVEN_SCORE || p.score == KVC_SCORE_SNAPKV_POOL) && p.pool_kernel > 2 * kMaxPoolHalo)
                return KVC_ERR_UNSUPPORTED;
        }
        if (C == 0) continue;
        if (!x.k_in || !x.v_in || !x.k_out || !x.v_out) return KVC_ERR_INVALID_ARG;
        if (C * cpr * 2 > 0x7fffffffLL) return KVC_ERR_TOO_LARGE;
        const uintptr_t a = (uintptr_t)x.k_in | (uintptr_t)x.v_in | (uintptr_t)x.k_out | (uintptr_t)x.v_out;
        if (a & 15) return KVC_ERR_UNSUPPORTED;
        const int64_t s = (x.k_stride_b | x.k_stride_h | x.k_stride_s | x.v_stride_b | x.v_stride_h | x.v_stride_s);
        if ((s * e) & 15) return KVC_ERR_UNSUPPORTED;
    }
    DeviceGuard guard(shape->device);
    NvtxRange nvtx_range("kvc_compress_layers_ws");
    if (guard.status != KVC_OK) return guard.status;

    for (int l0 = 0; l0 < n_layers; l0 += KVC_MAX_LAYERS_PER_LAUNCH) {
        const int nl = (n_layers - l0) < KVC_MAX_LAYERS_PER_LAUNCH ? (n_layers - l0) : KVC_MAX_LAYERS_PER_LAUNCH;
        BatchDev bd;
        memset(&bd, 0, sizeof(bd));
        bd.B = B;
        bd.H = H;
        bd.cpr = cpr;
        int n_active = 0, max_region = 0, max_ksel = 0;
        int64_t rows_moved = 0;
        bool any_select = false, any_scan = false;
        for (int l = 0; l < nl; ++l) {
            const kvc_layer_plan& p = plans[l0 + l];
            const kvc_layer_io& x = io[l0 + l];
            if (p.sink + p.k_sel + p.tail == 0) continue;
            rows_moved += (p.k_sel > 0 ? p.sel_hi - p.sel_lo : 0) + 4LL * (p.sink + p.k_sel + p.tail);
            LayerDev& d = bd.layers[n_active++];
            d.k_in = (const char*)x.k_in;
            d.v_in = (const char*)x.v_in;
            d.k_out = (char*)x.k_out;
            d.v_out = (char*)x.v_out;
            d.idx_out = x.idx_out;
            d.idx_in = p.score == KVC_SCORE_GIVEN_SCORE ? (const int32_t*)x.score_in : x.idx_in;
            d.ksb = x.k_stride_b * e;
            d.ksh = x.k_stride_h * e;
            d.kss = x.k_stride_s * e;
            d.vsb = x.v_stride_b * e;
            d.vsh = x.v_stride_h * e;
            d.vss = x.v_stride_s * e;
            d.S = p.seq_len;
            d.sink = p.sink;
            d.lo = p.sel_lo;
            d.hi = p.sel_hi;
            d.ksel = p.k_sel;
            d.tail = p.tail;
            d.score = p.k_sel > 0 ? p.score : KVC_SCORE_NONE;
            d.pool = p.pool_kernel;
            const bool ranked = p.k_sel > 0 && (p.score == KVC_SCORE_L2_LOW || p.score == KVC_SCORE_L2_HIGH ||
                                                p.score == KVC_SCORE_SNAPKV_POOL);
            if (ranked && x.norms_in) {  // the caller holds the key norms: the scan is skipped
                d.n_in = (const char*)x.norms_in;
                d.nsb = x.n_stride_b * key_bytes(dt);
                d.nsh = x.n_stride_h * key_bytes(dt);
            }
            if (p.k_sel > 0) {
                any_select = true;
                any_scan |= ranked && !x.norms_in;
                if (p.k_sel > max_ksel) max_ksel = p.k_sel;
                if (p.score != KVC_SCORE_GIVEN_INDEX && p.sel_hi - p.sel_lo > max_region)
                    max_region = p.sel_hi - p.sel_lo;
            }
        }
        if (n_active == 0) continue;
        bd.idx_cap = (max_ksel + 3) & ~3;
        // no K scan in this launch (stored norms, caller-supplied scores or rows, pure slices with a select next to
        // them): only kept rows move -> residency over slot depth
        const bool light = any_select && !any_scan;
        TmaPlan tp = plan_tma(dt, cpr, max_region, bd.idx_cap, any_select, light, (int64_t)B * H * n_active,
                              rows_moved / n_active * cpr * 16);
        if (any_select && workspace != nullptr && !onchip_plan_ok(tp, cpr, light)) {
            // keys and kept indices go to the workspace; shared memory keeps the histogram and the slots
            const WsLayout w = ws_layout(dt, max_region, bd.idx_cap);
            if (w.unit * B * H * n_active > workspace_bytes) return KVC_ERR_INVALID_ARG;
            bd.ws = (char*)workspace;
            bd.ws_unit = w.unit;
            bd.ws_keys = w.keys;
            tp = plan_tma(dt, cpr, 0, 0, true, light);
        }
        if (!tp.ok) return KVC_ERR_TOO_LARGE;
        FusedFn fn = pick_tma(dt, cpr, tp.nt);
        bd.nsw = tp.nsw;
        bd.off_hist = tp.off_hist;
        bd.off_idx = tp.off_idx;
        bd.off_keys = tp.off_keys;
        bd.off_stage = tp.off_stage;
        bd.upc = tp.upc;
        dim3 grid((unsigned)(((int64_t)B * H + tp.upc - 1) / tp.upc), (unsigned)n_active, 1);
        st = ensure_tma_attrs((const void*)fn, shape->device);
        if (st != KVC_OK) return st;
        fn<<<grid, tp.nt, tp.smem, (cudaStream_t)stream>>>(bd);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return cuda_fail(err, "kvc_fused_tma_kernel launch");
        g_launches.fetch_add(1);
    }
    return KVC_OK;
}

int kvc_key_norms(const kvc_shape* shape, const void* k_in, int64_t stride_b, int64_t stride_h, int64_t stride_s,
                  int32_t row_lo, int32_t row_hi, void* norms_out, void* stream) {
    if (!shape || !k_in || !norms_out || row_lo < 0 || row_hi < row_lo) return KVC_ERR_INVALID_ARG;
    const int B = shape->batch, H = shape->heads, D = shape->head_dim, dt = shape->dtype;
    if (B <= 0 || H <= 0 || D <= 0) return KVC_ERR_INVALID_ARG;
    if (dt != KVC_DTYPE_F32 && dt != KVC_DTYPE_F16 && dt != KVC_DTYPE_BF16) return KVC_ERR_UNSUPPORTED;
    const int e = elem_bytes(dt);
    if (((int64_t)D * e) % 16 != 0 || ((uintptr_t)k_in & 15)) return KVC_ERR_UNSUPPORTED;
    if (((stride_b | stride_h | stride_s) * e) & 15) return KVC_ERR_UNSUPPORTED;
    const int R = row_hi - row_lo;
    const int64_t total = (int64_t)B * H * R;
    if (total == 0) return KVC_OK;
    DeviceGuard guard(shape->device);
    NvtxRange nvtx_range("kvc_key_norms");
    if (guard.status != KVC_OK) return guard.status;
    const int cpr = D * e / 16;
    const int lpr = pow2_lanes(cpr), cpl = cpr / lpr;
    const int rpw = 32 / lpr;
    const int64_t warps_needed = (total + rpw - 1) / rpw;
    int64_t blocks = (warps_needed + 7) / 8;
    const int64_t cap = 148LL * 8 * 4;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t sb = stride_b * e, sh = stride_h * e, ss = stride_s * e;
    switch (dt) {
        case KVC_DTYPE_F32:
            kvc_norm_kernel<KVC_DTYPE_F32><<<(unsigned)blocks, 256, 0, s>>>((const char*)k_in, sb, sh, ss, H, row_lo, R,
                                                                           total, lpr, cpl, (uint32_t*)norms_out);
            break;
        case KVC_DTYPE_F16:
            kvc_norm_kernel<KVC_DTYPE_F16><<<(unsigned)blocks, 256, 0, s>>>((const char*)k_in, sb, sh, ss, H, row_lo, R,
                                                                           total, lpr, cpl, (uint16_t*)norms_out);
            break;
        default:
            kvc_norm_kernel<KVC_DTYPE_BF16><<<(unsigned)blocks, 256, 0, s>>>((const char*)k_in, sb, sh, ss, H, row_lo,
                                                                            R, total, lpr, cpl, (uint16_t*)norms_out);
            break;
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return cuda_fail(err, "kvc_norm_kernel launch");
    g_launches.fetch_add(1);
    return KVC_OK;
}

int kvc_select(int32_t dtype, int32_t device, const void* scores, int64_t n_rows, int32_t n, int32_t k,
               int32_t largest, int32_t* idx_out, void* stream) {
    if (!scores || !idx_out || n_rows < 0 || n < 0 || k < 0 || k > n) return KVC_ERR_INVALID_ARG;
    if (dtype != KVC_DTYPE_F32 && dtype != KVC_DTYPE_F16 && dtype != KVC_DTYPE_BF16) return KVC_ERR_UNSUPPORTED;
    if (n_rows == 0 || k == 0) return KVC_OK;
    if (n_rows > 0x7fffffffLL) return KVC_ERR_INVALID_ARG;
    const size_t smem = fused_smem_bytes(dtype, n, (k + 3) & ~3);
    if (smem > (size_t)kMaxSmemOptin) return KVC_ERR_TOO_LARGE;
    DeviceGuard guard(device);
    NvtxRange nvtx_range("kvc_select");
    if (guard.status != KVC_OK) return guard.status;
    int st = KVC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const void* fn = nullptr;
    switch (dtype) {
        case KVC_DTYPE_F32: fn = (const void*)kvc_select_kernel<KVC_DTYPE_F32, kNT>; break;
        case KVC_DTYPE_F16: fn = (const void*)kvc_select_kernel<KVC_DTYPE_F16, kNT>; break;
        default: fn = (const void*)kvc_select_kernel<KVC_DTYPE_BF16, kNT>; break;
    }
    st = ensure_smem(fn, smem);
    if (st != KVC_OK) return st;
    switch (dtype) {
        case KVC_DTYPE_F32:
            kvc_select_kernel<KVC_DTYPE_F32, kNT>
                <<<(unsigned)n_rows, kNT, smem, s>>>((const uint32_t*)scores, n, k, largest, idx_out);
            break;
        case KVC_DTYPE_F16:
            kvc_select_kernel<KVC_DTYPE_F16, kNT>
                <<<(unsigned)n_rows, kNT, smem, s>>>((const uint16_t*)scores, n, k, largest, idx_out);
            break;
        default:
            kvc_select_kernel<KVC_DTYPE_BF16, kNT>
                <<<(unsigned)n_rows, kNT, smem, s>>>((const uint16_t*)scores, n, k, largest, idx_out);
            break;
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return cuda_fail(err, "kvc_select_kernel launch");
    g_launches.fetch_add(1);
    return KVC_OK;
}

#endif  // KVC_HAS_CORE

#if KVC_HAS_SLAB
int kvc_slab_append(const kvc_shape* shape, int32_t n_layers, const kvc_slab_layer* slabs,
                    const kvc_slab_new_rows* rows, void* stream) {
    int cpr = 0;
    int st = check_shape(shape, &cpr);
    if (st != KVC_OK) return st;
    if (n_layers < 0 || (n_layers > 0 && (!slabs || !rows))) return KVC_ERR_INVALID_ARG;
    if (n_layers == 0) return KVC_OK;
    if (!row_width_ok(cpr)) return KVC_ERR_UNSUPPORTED;
    const int B = shape->batch, H = shape->heads, dt = shape->dtype;
    const int e = elem_bytes(dt);
    for (int l = 0; l < n_layers; ++l) {
        const kvc_slab_layer& sl = slabs[l];
        const kvc_slab_new_rows& r = rows[l];
        if (r.n_new < 0 || r.cur_len < 0) return KVC_ERR_INVALID_ARG;
        if (r.n_new == 0) continue;
        if (!sl.k || !sl.v || !sl.norms || !r.k_new || !r.v_new) return KVC_ERR_INVALID_ARG;
        const uintptr_t a = (uintptr_t)sl.k | (uintptr_t)sl.v | (uintptr_t)r.k_new | (uintptr_t)r.v_new;
        if (a & 15) return KVC_ERR_UNSUPPORTED;
        const int64_t sb = sl.k_stride_b | sl.k_stride_h | sl.v_stride_b | sl.v_stride_h | r.k_stride_b | r.k_stride_h |
                           r.k_stride_s | r.v_stride_b | r.v_stride_h | r.v_stride_s;
        if ((sb * e) & 15) return KVC_ERR_UNSUPPORTED;
        if ((int64_t)B * H * r.n_new > 0x7fffffffLL) return KVC_ERR_TOO_LARGE;
    }
    DeviceGuard guard(shape->device);
    NvtxRange nvtx_range("kvc_slab_append");
    if (guard.status != KVC_OK) return guard.status;
    auto fill = [&](AppendLayerDev& d, const kvc_slab_layer& sl, const kvc_slab_new_rows& r) {
        d.k_new = (const char*)r.k_new;
        d.v_new = (const char*)r.v_new;
        d.k = (char*)sl.k;
        d.v = (char*)sl.v;
        d.n = (char*)sl.norms;
        d.nksb = r.k_stride_b * e;
        d.nksh = r.k_stride_h * e;
        d.nkss = r.k_stride_s * e;
        d.nvsb = r.v_stride_b * e;
        d.nvsh = r.v_stride_h * e;
        d.nvss = r.v_stride_s * e;
        d.ksb = sl.k_stride_b * e;
        d.ksh = sl.k_stride_h * e;
        d.vsb = sl.v_stride_b * e;
        d.vsh = sl.v_stride_h * e;
        d.nsb = sl.n_stride_b * key_bytes(dt);
        d.nsh = sl.n_stride_h * key_bytes(dt);
        d.cur_len = r.cur_len;
        d.n_new = r.n_new;
    };
    if (n_layers == 1 && rows[0].n_new > 0 && rows[0].n_new < 32) {
        // the per-layer update of a decode loop: small parameter block, one small launch
        AppendOneDev one;
        one.B = B;
        one.H = H;
        one.max_new = rows[0].n_new;
        one.cpr = cpr;
        fill(one.layers[0], slabs[0], rows[0]);
        const int64_t threads = (int64_t)B * H * rows[0].n_new;
        pick_append_one(dt, cpr)<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(one);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return cuda_fail(err, "kvc_slab_append_kernel launch");
        g_launches.fetch_add(1);
        return KVC_OK;
    }
    AppendFn fn = pick_append(dt, cpr);
    for (int l0 = 0; l0 < n_layers; l0 += KVC_MAX_LAYERS_PER_LAUNCH) {
        const int nl = (n_layers - l0) < KVC_MAX_LAYERS_PER_LAUNCH ? (n_layers - l0) : KVC_MAX_LAYERS_PER_LAUNCH;
        AppendBatchDev bd;
        memset(&bd, 0, sizeof(bd));
        bd.B = B;
        bd.H = H;
        int n_active = 0, max_new = 0;
        for (int l = 0; l < nl; ++l) {
            const kvc_slab_layer& sl = slabs[l0 + l];
            const kvc_slab_new_rows& r = rows[l0 + l];
            if (r.n_new == 0) continue;
            fill(bd.layers[n_active++], sl, r);
            if (r.n_new > max_new) max_new = r.n_new;
        }
        if (n_active == 0) continue;
        bd.max_new = max_new;
        bd.cpr = cpr;
        cudaError_t err;
        if (max_new >= 32) {
            // prefill / chunked prefill: rows move 32 at a time through shared memory with bulk copies
            AppendFn tfn = pick_append_tma(dt, cpr);
            st = ensure_tma_attrs((const void*)tfn, shape->device);
            if (st != KVC_OK) return st;
            const size_t smem = 128 + (size_t)8 * 32 * cpr * 16;
            const int64_t items = (int64_t)B * H * ((max_new + 31) / 32);
            int64_t blocks = (items + 7) / 8;
            const int64_t cap = 148LL * 3 * 8;  // a few waves; warps stride over the remaining blocks
            if (blocks > cap) blocks = cap;
            dim3 grid((unsigned)blocks, (unsigned)n_active, 1);
            tfn<<<grid, 256, smem, (cudaStream_t)stream>>>(bd);
            err = cudaGetLastError();
        } else {
            const int64_t threads = (int64_t)B * H * max_new;
            dim3 grid((unsigned)((threads + 127) / 128), (unsigned)n_active, 1);
            fn<<<grid, 128, 0, (cudaStream_t)stream>>>(bd);
            err = cudaGetLastError();
        }
        if (err != cudaSuccess) return cuda_fail(err, "kvc_slab_append_kernel launch");
        g_launches.fetch_add(1);
    }
    return KVC_OK;
}

int kvc_slab_compress(const kvc_shape* shape, int32_t n_layers, const kvc_layer_plan* plans,
                      const kvc_slab_layer* slabs, int32_t* const* idx_out, const int32_t* const* idx_in,
                      void* workspace, int64_t workspace_bytes, void* stream) {
    int cpr = 0;
    int st = check_shape(shape, &cpr);
    if (st != KVC_OK) return st;
    if (n_layers < 0 || (n_layers > 0 && (!slabs || !plans))) return KVC_ERR_INVALID_ARG;
    if (n_layers == 0) return KVC_OK;
    if (!row_width_ok(cpr)) return KVC_ERR_UNSUPPORTED;
    const int B = shape->batch, H = shape->heads, dt = shape->dtype;
    const int e = elem_bytes(dt);
    for (int l = 0; l < n_layers; ++l) {
        const kvc_layer_plan& p = plans[l];
        const kvc_slab_layer& sl = slabs[l];
        if (p.seq_len < 0 || p.sink < 0 || p.k_sel < 0 || p.tail < 0) return KVC_ERR_INVALID_ARG;
        if (p.sink > p.seq_len || p.tail > p.seq_len) return KVC_ERR_INVALID_ARG;
        if ((int64_t)p.sink + p.k_sel + p.tail > p.seq_len) return KVC_ERR_INVALID_ARG;
        if (p.k_sel > 0) {
            if (p.sel_lo < p.sink || p.sel_hi > p.seq_len - p.tail || p.sel_lo > p.sel_hi) return KVC_ERR_INVALID_ARG;
            if (p.k_sel > p.sel_hi - p.sel_lo) return KVC_ERR_INVALID_ARG;
            if (p.score <= KVC_SCORE_NONE || p.score > KVC_SCORE_GIVEN_SCORE) return KVC_ERR_INVALID_ARG;
            if ((p.score == KVC_SCORE_GIVEN_INDEX || p.score == KVC_SCORE_GIVEN_SCORE) && (!idx_in || !idx_in[l]))
                return KVC_ERR_INVALID_ARG;
            if ((p.score == KVC_SCORE_SNAPKV_POOL || p.score == KVC_SCORE_GIVEN_SCORE) && p.pool_kernel > 2 * kMaxPoolHalo)
                return KVC_ERR_UNSUPPORTED;
        }
        if (p.sink + p.k_sel + p.tail == 0) continue;
        if (!sl.k || !sl.v || !sl.norms) return KVC_ERR_INVALID_ARG;
        if (((uintptr_t)sl.k | (uintptr_t)sl.v) & 15) return KVC_ERR_UNSUPPORTED;
        if (((sl.k_stride_b | sl.k_stride_h | sl.v_stride_b | sl.v_stride_h) * e) & 15) return KVC_ERR_UNSUPPORTED;
    }
    DeviceGuard guard(shape->device);
    NvtxRange nvtx_range("kvc_slab_compress");
    if (guard.status != KVC_OK) return guard.status;
    for (int l0 = 0; l0 < n_layers; l0 += KVC_MAX_LAYERS_PER_LAUNCH) {
        const int nl = (n_layers - l0) < KVC_MAX_LAYERS_PER_LAUNCH ? (n_layers - l0) : KVC_MAX_LAYERS_PER_LAUNCH;
        SlabBatchDev bd;
        memset(&bd, 0, sizeof(bd));
        bd.B = B;
        bd.H = H;
        bd.cpr = cpr;
        int n_active = 0, max_region = 0, max_ksel = 0;
        bool any_select = false;
        for (int l = 0; l < nl; ++l) {
            const kvc_layer_plan& p = plans[l0 + l];
            const kvc_slab_layer& sl = slabs[l0 + l];
            if (p.sink + p.k_sel + p.tail == 0) continue;
            SlabLayerDev& d = bd.layers[n_active++];
            d.k = (char*)sl.k;
            d.v = (char*)sl.v;
            d.n = (char*)sl.norms;
            d.idx_out = idx_out ? idx_out[l0 + l] : nullptr;
            d.idx_in = idx_in ? idx_in[l0 + l] : nullptr;
            d.ksb = sl.k_stride_b * e;
            d.ksh = sl.k_stride_h * e;
            d.vsb = sl.v_stride_b * e;
            d.vsh = sl.v_stride_h * e;
            d.nsb = sl.n_stride_b * key_bytes(dt);
            d.nsh = sl.n_stride_h * key_bytes(dt);
            d.S = p.seq_len;
            d.sink = p.sink;
            d.lo = p.sel_lo;
            d.hi = p.sel_hi;
            d.ksel = p.k_sel;
            d.tail = p.tail;
            d.score = p.k_sel > 0 ? p.score : KVC_SCORE_NONE;
            d.pool = p.pool_kernel;
            if (p.k_sel > 0) {
                any_select = true;
                if (p.k_sel > max_ksel) max_ksel = p.k_sel;
                if (p.score != KVC_SCORE_GIVEN_INDEX && p.sel_hi - p.sel_lo > max_region) max_region = p.sel_hi - p.sel_lo;
            }
        }
        if (n_active == 0) continue;
        bd.idx_cap = (max_ksel + 3) & ~3;
        TmaPlan tp = plan_tma(dt, cpr, max_region, bd.idx_cap, any_select, /*light_traffic=*/true, 0, 0, /*in_place=*/true);
        if (any_select && workspace != nullptr && !onchip_plan_ok(tp, cpr, true)) {
            const WsLayout w = ws_layout(dt, max_region, bd.idx_cap);
            if (w.unit * B * H * n_active > workspace_bytes) return KVC_ERR_INVALID_ARG;
            bd.ws = (char*)workspace;
            bd.ws_unit = w.unit;
            bd.ws_keys = w.keys;
            tp = plan_tma(dt, cpr, 0, 0, true, /*light_traffic=*/true);
        }
        if (!tp.ok) return KVC_ERR_TOO_LARGE;
        SlabFn fn = pick_slab(dt, cpr, tp.nt);
        bd.nsw = tp.nsw;
        bd.off_hist = tp.off_hist;
        bd.off_idx = tp.off_idx;
        bd.off_keys = tp.off_keys;
        bd.off_stage = tp.off_stage;
        st = ensure_tma_attrs((const void*)fn, shape->device);
        if (st != KVC_OK) return st;
        dim3 grid((unsigned)((int64_t)B * H), (unsigned)n_active, 1);
        fn<<<grid, tp.nt, tp.smem, (cudaStream_t)stream>>>(bd);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return cuda_fail(err, "kvc_slab_compress_kernel launch");
        g_launches.fetch_add(1);
    }
    return KVC_OK;
}

#endif  // KVC_HAS_SLAB

#if KVC_HAS_VOTE
// cuTensorMapEncodeTiled through the runtime (the library links cudart statically and never links libcuda).
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// Shared-memory bytes the fused tail needs inside the (dead) key-tile ring: histogram, kept indices, radix keys
// of the prefix, at least one staging slot of 32 rows.
static size_t vote_tail_bytes(int cpr, int prefix_rows, int idx_cap) {
    const size_t keys = ((size_t)prefix_rows * 2 + 15) & ~(size_t)15;
    return (((size_t)kHistBins * 4 + (size_t)idx_cap * 4 + keys + 127) & ~(size_t)127) + (size_t)32 * cpr * 16;
}

// TMA-fed vote kernel (kvc_vote.cuh): returns KVC_ERR_UNSUPPORTED when a tensor map cannot describe the keys.
// plans / io non-null: the fused form — pool, select and gather behind the vote in the same launch.
static int launch_vote_tma(const kvc_shape* shape, int32_t n_layers, const kvc_vote_layer* layers, int32_t group,
                           int32_t window, int cpr, const kvc_layer_plan* plans, const kvc_layer_io* io, void* stream) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return KVC_ERR_UNSUPPORTED;
    const int B = shape->batch, H = shape->heads, D = shape->head_dim, dt = shape->dtype;
    using Fn = void (*)(const VoteTmaBatchDev);
    Fn fn = nullptr;
    if (dt == KVC_DTYPE_BF16)
        fn = cpr == 8 ? kvc_snapkv_vote_tma_kernel<KVC_DTYPE_BF16, 8>
                      : cpr == 10 ? kvc_snapkv_vote_tma_kernel<KVC_DTYPE_BF16, 10> : kvc_snapkv_vote_tma_kernel<KVC_DTYPE_BF16, 16>;
    else
        fn = cpr == 8 ? kvc_snapkv_vote_tma_kernel<KVC_DTYPE_F16, 8>
                      : cpr == 10 ? kvc_snapkv_vote_tma_kernel<KVC_DTYPE_F16, 10> : kvc_snapkv_vote_tma_kernel<KVC_DTYPE_F16, 16>;
    const size_t tile = (size_t)(cpr / 8) * kVoteTile * 128 + (size_t)(cpr % 8) * kVoteTile * 16;
    const size_t smem = 6144 + (size_t)(kVoteM / 8) * cpr * kVoteLBO + (size_t)vote_tma_ring(cpr) * tile;
    int st = ensure_tma_attrs((const void*)fn, shape->device);
    if (st != KVC_OK) return st;
    for (int l0 = 0; l0 < n_layers; l0 += 32) {
        const int nl = (n_layers - l0) < 32 ? (n_layers - l0) : 32;
        static_assert(sizeof(VoteTmaBatchDev) < 32 * 1024, "kernel parameters are limited to 32 KB");
        VoteTmaBatchDev bd;
        memset(&bd, 0, sizeof(bd));
        bd.B = B;
        bd.H = H;
        bd.G = group;
        bd.W = window;
        bd.scale_log2e = 1.4426950408889634f / sqrtf((float)D);
        bd.pad[0] = vote_debug_mode();
#ifdef KVC_LAB
        {   // lab: KVC_VOTE_TIMELINE=<hex device pointer> of a [CTAs][8] int64 buffer
            const char* tl = getenv("KVC_VOTE_TIMELINE");
            bd.lab_timeline = (tl && *tl) ? (long long*)strtoull(tl, nullptr, 16) : nullptr;
        }
#endif
        for (int l = 0; l < nl; ++l) {
            const kvc_vote_layer& v = layers[l0 + l];
            VoteTmaLayerDev& d = bd.layers[l];
            const cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)v.seq_len, (cuuint64_t)H, (cuuint64_t)B};
            const cuuint64_t strides[3] = {(cuuint64_t)v.k_stride_s * 2, (cuuint64_t)v.k_stride_h * 2,
                                           (cuuint64_t)v.k_stride_b * 2};
            const cuuint32_t box[4] = {64, (cuuint32_t)kVoteTile, 1, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            const CUresult r = encode(&d.map, dt == KVC_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                      4, const_cast<void*>(v.k_in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return KVC_ERR_UNSUPPORTED;
            if (cpr % 8) {  // D = 80: the last 16 elements of every row through a 32-byte-swizzled box
                const cuuint32_t box_tail[4] = {16, (cuuint32_t)kVoteTile, 1, 1};
                const CUresult r2 = encode(&d.map_tail, dt == KVC_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                           4, const_cast<void*>(v.k_in), dims, strides, box_tail, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r2 != CUDA_SUCCESS) return KVC_ERR_UNSUPPORTED;
            }
            d.q = (const char*)v.q_obs;
            d.votes = (char*)v.votes_out;
            d.lse = v.lse;
            d.qsb = v.q_stride_b * 2;
            d.qsh = v.q_stride_h * 2;
            d.qss = v.q_stride_s * 2;
            d.S = v.seq_len;
            if (plans != nullptr) {
                const kvc_layer_plan& p = plans[l0 + l];
                const kvc_layer_io& x = io[l0 + l];
                d.k_in = (const char*)x.k_in;
                d.v_in = (const char*)x.v_in;
                d.k_out = (char*)x.k_out;
                d.v_out = (char*)x.v_out;
                d.idx_out = x.idx_out;
                d.ksb = x.k_stride_b * 2;
                d.ksh = x.k_stride_h * 2;
                d.kss = x.k_stride_s * 2;
                d.vsb = x.v_stride_b * 2;
                d.vsh = x.v_stride_h * 2;
                d.vss = x.v_stride_s * 2;
                d.ksel = p.k_sel;
                d.tail = p.tail;
                d.pool = p.pool_kernel;
                d.idx_cap = (p.k_sel + 3) & ~3;
            }
        }
        dim3 grid((unsigned)((int64_t)B * H), (unsigned)nl, 1);
        fn<<<grid, 576, smem, (cudaStream_t)stream>>>(bd);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return cuda_fail(err, "kvc_snapkv_vote_tma_kernel launch");
        g_launches.fetch_add(1);
    }
    return KVC_OK;
}

int kvc_snapkv_vote(const kvc_shape* shape, int32_t n_layers, const kvc_vote_layer* layers, int32_t group,
                    int32_t window, void* stream) {
    int cpr = 0;
    int st = check_shape(shape, &cpr);
    if (st != KVC_OK) return st;
    if (n_layers < 0 || (n_layers > 0 && !layers) || group <= 0 || window <= 0) return KVC_ERR_INVALID_ARG;
    if (n_layers == 0) return KVC_OK;
    if (shape->dtype == KVC_DTYPE_F32) return KVC_ERR_UNSUPPORTED;
    if (cpr != 8 && cpr != 10 && cpr != 16) return KVC_ERR_UNSUPPORTED;
    if ((int64_t)group * window > kVoteM) return KVC_ERR_UNSUPPORTED;
    for (int l = 0; l < n_layers; ++l) {
        const kvc_vote_layer& v = layers[l];
        if (!v.k_in || !v.q_obs || !v.votes_out || v.seq_len <= window) return KVC_ERR_INVALID_ARG;
        if (((uintptr_t)v.k_in | (uintptr_t)v.q_obs) & 15) return KVC_ERR_UNSUPPORTED;
        const int64_t sb = v.k_stride_b | v.k_stride_h | v.k_stride_s | v.q_stride_b | v.q_stride_h | v.q_stride_s;
        if ((sb * 2) & 15) return KVC_ERR_UNSUPPORTED;
    }
    DeviceGuard guard(shape->device);
    NvtxRange nvtx_range("kvc_snapkv_vote");
    if (guard.status != KVC_OK) return guard.status;
    // key tiles arrive through TMA tensor loads: layouts a tensor map cannot describe are KVC_ERR_UNSUPPORTED (the
    // Python layer makes such keys contiguous first)
    return launch_vote_tma(shape, n_layers, layers, group, window, cpr, nullptr, nullptr, stream);
}

int kvc_snapkv_vote_compress(const kvc_shape* shape, int32_t n_layers, const kvc_vote_layer* layers,
                             const kvc_layer_plan* plans, const kvc_layer_io* io, int32_t group, int32_t window,
                             void* stream) {
    int cpr = 0;
    int st = check_shape(shape, &cpr);
    if (st != KVC_OK) return st;
    if (n_layers < 0 || (n_layers > 0 && (!layers || !plans || !io)) || group <= 0 || window <= 0)
        return KVC_ERR_INVALID_ARG;
    if (n_layers == 0) return KVC_OK;
    if (shape->dtype == KVC_DTYPE_F32) return KVC_ERR_UNSUPPORTED;
    if (cpr != 8 && cpr != 10 && cpr != 16) return KVC_ERR_UNSUPPORTED;
    if ((int64_t)group * window > kVoteM) return KVC_ERR_UNSUPPORTED;
    const size_t ring_bytes = (size_t)vote_tma_ring(cpr) * ((size_t)(cpr / 8) * kVoteTile * 128 + (size_t)(cpr % 8) * kVoteTile * 16);
    for (int l = 0; l < n_layers; ++l) {
        const kvc_vote_layer& v = layers[l];
        const kvc_layer_plan& p = plans[l];
        const kvc_layer_io& x = io[l];
        if (!v.k_in || !v.q_obs || !v.votes_out || v.seq_len <= window) return KVC_ERR_INVALID_ARG;
        if (((uintptr_t)v.k_in | (uintptr_t)v.q_obs) & 15) return KVC_ERR_UNSUPPORTED;
        const int64_t sb = v.k_stride_b | v.k_stride_h | v.k_stride_s | v.q_stride_b | v.q_stride_h | v.q_stride_s;
        if ((sb * 2) & 15) return KVC_ERR_UNSUPPORTED;
        // the plan of snapkv_lite in vote mode: no sinks, the prefix [0, S - W) ranked by pooled votes, the window kept
        if (p.seq_len != v.seq_len || p.sink != 0 || p.sel_lo != 0 || p.sel_hi != v.seq_len - window || p.k_sel <= 0 ||
            p.k_sel > p.sel_hi || p.tail < 0 || p.tail > window || p.score != KVC_SCORE_GIVEN_SCORE)
            return KVC_ERR_INVALID_ARG;
        if (p.pool_kernel > 2 * kMaxPoolHalo) return KVC_ERR_UNSUPPORTED;
        if (!x.k_in || !x.v_in || !x.k_out || !x.v_out || x.k_in != v.k_in) return KVC_ERR_INVALID_ARG;
        if (((uintptr_t)x.v_in | (uintptr_t)x.k_out | (uintptr_t)x.v_out) & 15) return KVC_ERR_UNSUPPORTED;
        if (((x.v_stride_b | x.v_stride_h | x.v_stride_s) * 2) & 15) return KVC_ERR_UNSUPPORTED;
        if (x.k_stride_b != v.k_stride_b || x.k_stride_h != v.k_stride_h || x.k_stride_s != v.k_stride_s)
            return KVC_ERR_INVALID_ARG;
        if (vote_tail_bytes(cpr, p.sel_hi, (p.k_sel + 3) & ~3) > ring_bytes) return KVC_ERR_TOO_LARGE;
    }
    DeviceGuard guard(shape->device);
    NvtxRange nvtx_range("kvc_snapkv_vote_compress");
    if (guard.status != KVC_OK) return guard.status;
    return launch_vote_tma(shape, n_layers, layers, group, window, cpr, plans, io, stream);
}

#endif  // KVC_HAS_VOTE

}  // extern "C"
