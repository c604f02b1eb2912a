// kvc_fused_tma.cuh — the bulk-copy (TMA) form of the fused compress kernel for sm_100a.
//
// One CTA owns one (layer, batch, head) unit, as in the LDG form, but no key/value byte passes
// through registers on its way in or out:
//
//   scan   each staging warp owns one shared-memory slot of 32 rows and one mbarrier.  Lane 0
//          streams the unit's key rows into the slot with `cp.async.bulk` (one copy per 32 rows;
//          one per row when rows are not contiguous), the warp waits on the mbarrier, and every
//          lane reduces ONE whole row from shared memory: 128-bit LDS, fp32 FMA per 16-byte
//          chunk, chunk sums combined by a balanced tree.  Lanes read their chunks in an
//          XOR-swizzled order so that the 8 lanes of an LDS.128 phase hit 8 different bank
//          groups although the row pitch (D*e bytes) is a multiple of 32 B; the tree makes the
//          sum independent of that order (a+b == b+a), so a row's norm does not depend on the
//          lane that computed it.  (torch.norm, e.g. l2_compress.py:70)
//   select unchanged: radix select over the on-chip keys (kvc_device.cuh).
//   gather output rows are produced 32 at a time per staging warp: every run of consecutive
//          source rows becomes one bulk load into the slot, and the 32 rows leave with ONE bulk
//          store into the dense output (expand + gather x2 + cat x2, e.g. fix_size_l2.py:132-147).
//
// Bytes in flight per SM = (resident CTAs) x (staging warps) x 32 x D*e, chosen by the host.
#pragma once
#include "kvc_device.cuh"
#include "kvc_tma.cuh"

namespace kvc {

// log2 of the XOR-swizzle group: min(ctz(CPR), 3)
__host__ __device__ constexpr int swz_bits(int cpr) { return (cpr % 8 == 0) ? 3 : (cpr % 4 == 0) ? 2 : (cpr % 2 == 0) ? 1 : 0; }

// Sum of squares of one row held in a staging slot.  `row_addr` = shared address of the row.
template <int DT, int CPR>
__device__ __forceinline__ float row_sumsq_smem(uint32_t row_addr, int lane) {
    using Tr = Traits<DT>;
    constexpr int SB = swz_bits(CPR);
    constexpr int G = 1 << SB;
    const int x = (lane >> (3 - SB)) & (G - 1);
    uint32_t base[G];
#pragma unroll
    for (int m = 0; m < G; ++m) base[m] = row_addr + (uint32_t)(((m ^ x) - m) * 16);
    float acc[CPR];
#pragma unroll
    for (int c = 0; c < CPR; ++c) acc[c] = Tr::sumsq(lds128(base[c & (G - 1)] + c * 16), 0.f);
    // balanced tree inside each aligned group of G chunks (invariant under the XOR), then groups in order
    float tot = 0.f;
#pragma unroll
    for (int g = 0; g < CPR / G; ++g) {
#pragma unroll
        for (int s = 1; s < G; s <<= 1) {
#pragma unroll
            for (int m = 0; m < G; m += 2 * s) acc[g * G + m] += acc[g * G + m + s];
        }
        tot = (g == 0) ? acc[0] : tot + acc[g * G];
    }
    return tot;
}

// Load `rows` (1..32) rows into the warp's slot; lane l supplies the global address of row l.
// Consecutive rows (when `dense`: row stride == row bytes) are merged into one bulk copy per run.
// All 32 lanes must call; completion is observed with mbar_wait(bar, parity).
template <int RB>
__device__ __forceinline__ void warp_load_rows(uint32_t slot, uint32_t bar, const char* src, int rows, bool dense,
                                               int lane) {
    const bool in = lane < rows;
    const unsigned long long a = (unsigned long long)src;
    const unsigned long long prev = __shfl_up_sync(0xffffffffu, a, 1);
    const bool head = in && (lane == 0 || !dense || a != prev + RB);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)rows * RB);
    __syncwarp();
    if (head) {
        const uint32_t later = heads & ~((2u << lane) - 1u);  // heads above this lane
        const int next = later ? (__ffs(later) - 1) : rows;
        bulk_g2s(slot + (uint32_t)lane * RB, src, (uint32_t)(next - lane) * RB, bar);
    }
}

template <int DT, int CPR, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) kvc_fused_tma_kernel(const __grid_constant__ BatchDev bd) {
    using Tr = Traits<DT>;
    using Key = typename Tr::Key;
    constexpr int RB = CPR * 16;  // row bytes
    constexpr int kShift0 = Tr::kKeyBits - kHistBits;

    const LayerDev& L = bd.layers[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(128) unsigned char smem[];
    int32_t* misc = reinterpret_cast<int32_t*>(smem);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + bd.off_hist);
    int32_t* sidx = reinterpret_cast<int32_t*>(smem + bd.off_idx);
    Key* keys = reinterpret_cast<Key*>(smem + bd.off_keys);
    const int nsw = bd.nsw;
    const bool stager = warp < nsw;
    const uint32_t slot = smem_u32(smem + bd.off_stage) + (uint32_t)warp * (32 * RB);
    const uint32_t bar = smem_u32(smem + kMiscInts * 4) + (uint32_t)warp * 8;
    uint32_t parity = 0;

    const int R = L.hi - L.lo;
    const int ksel = L.ksel;
    const int score = L.score;
    const bool kdense = L.kss == RB, vdense = L.vss == RB;
    const bool scan = ksel > 0 && score != KVC_SCORE_GIVEN_INDEX && score != KVC_SCORE_GIVEN_SCORE;

    if (stager) {
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_init_fence();
        }
        __syncwarp();
    }
    // A CTA walks `upc` consecutive (batch, head) units: neighbours in memory, so the SM keeps
    // touching the same 2 MB pages (matters when only a sparse slice of each unit is read).
    const int n_units = bd.B * bd.H;
    const int bh_end = min(n_units, ((int)blockIdx.x + 1) * bd.upc);
    for (int bh = (int)blockIdx.x * bd.upc; bh < bh_end; ++bh) {
    const int b = bh / bd.H, h = bh - b * bd.H;
    if (bd.ws != nullptr) {  // selection too large for shared memory: keys / kept indices live in the workspace
        char* unit = bd.ws + ((int64_t)blockIdx.y * n_units + bh) * bd.ws_unit;
        keys = reinterpret_cast<Key*>(unit);
        sidx = reinterpret_cast<int32_t*>(unit + bd.ws_keys);
    }
    const char* kbase = L.k_in + (int64_t)b * L.ksb + (int64_t)h * L.ksh;
    const char* vbase = L.v_in + (int64_t)b * L.vsb + (int64_t)h * L.vsh;
    const char* kreg = kbase + (int64_t)L.lo * L.kss;
    const int nblk = (R + 31) >> 5;
    if (scan && stager && warp < nblk)  // first block is in flight while the histogram is cleared
        warp_load_rows<RB>(slot, bar, kreg + (int64_t)(warp * 32 + lane) * L.kss, min(32, R - warp * 32), kdense, lane);

    if (ksel > 0 && score == KVC_SCORE_GIVEN_INDEX) {
        const int32_t* src = L.idx_in + (int64_t)bh * ksel;
        for (int i = tid; i < ksel; i += NT) sidx[i] = src[i];
    } else if (ksel > 0 && score == KVC_SCORE_GIVEN_SCORE) {
        // ---------------------------------------------------------- caller-supplied scores: pool -> keep the highest
        for (int i = tid; i < kHistBins; i += NT) hist[i] = 0;
        const Key* src = reinterpret_cast<const Key*>(L.idx_in) + (int64_t)bh * R;
        load_keys_vectorised<DT, NT>(src, R, keys, [&](int i, uint32_t raw) { keys[i] = (Key)raw; });
        __syncthreads();
        snapkv_transform<DT, NT>(keys, R, L.pool, hist, misc, /*invert=*/false);
        block_radix_select<Key, NT>(keys, R, ksel, hist, misc, sidx, L.lo);
    } else if (ksel > 0) {
        // ---------------------------------------------------------- K1: scan
        for (int i = tid; i < kHistBins; i += NT) hist[i] = 0;
        if (tid == 0) misc[kMiscMaxRaw] = 0;
        __syncthreads();
        const bool snap = (score == KVC_SCORE_SNAPKV_POOL);
        const bool desc = (score == KVC_SCORE_L2_HIGH);
        uint32_t local_max = 0;
        if (stager) {
            for (int blk = warp; blk < nblk; blk += nsw) {
                const int r0 = blk << 5;
                const int rows = min(32, R - r0);
                mbar_wait(bar, parity);
                parity ^= 1;
                float ss = 0.f;
                if (lane < rows) ss = row_sumsq_smem<DT, CPR>(slot + (uint32_t)lane * RB, lane);
                __syncwarp();  // every lane has consumed its row: the slot may be refilled
                const int nb = blk + nsw;
                if (nb < nblk)
                    warp_load_rows<RB>(slot, bar, kreg + (int64_t)(nb * 32 + lane) * L.kss, min(32, R - nb * 32), kdense,
                                       lane);
                if (lane < rows) {
                    const uint32_t raw = Tr::to_raw(sqrtf(ss));
                    if (snap) {
                        keys[r0 + lane] = (Key)raw;
                        local_max = max(local_max, raw);
                    } else {
                        const Key key = ordered_key<Key>(raw, desc);
                        keys[r0 + lane] = key;
                        atomicAdd(&hist[(uint32_t)key >> kShift0], 1u);
                    }
                }
            }
        }
        if (snap) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
            if (lane == 0) atomicMax(reinterpret_cast<uint32_t*>(&misc[kMiscMaxRaw]), local_max);
        }
        __syncthreads();
        if (snap) snapkv_transform<DT, NT>(keys, R, L.pool, hist, misc);
        // ---------------------------------------------------------- K2: select
        block_radix_select<Key, NT>(keys, R, ksel, hist, misc, sidx, L.lo);
    }
    __syncthreads();  // sidx complete (GIVEN_INDEX path); no-op cost otherwise

    // -------------------------------------------------------------- K3: gather
    const int sink = L.sink;
    const int C = sink + ksel + L.tail;
    const int tail0 = L.S - L.tail - sink - ksel;  // src row = j + tail0 for tail rows
    if (L.idx_out != nullptr) {
        int32_t* io = L.idx_out + (int64_t)bh * C;
        for (int j = tid; j < C; j += NT) io[j] = j < sink ? j : (j < sink + ksel ? sidx[j - sink] : j + tail0);
    }
    if (stager) {
        const int nbc = (C + 31) >> 5;
        char* ko = L.k_out + (int64_t)bh * C * RB;
        char* vo = L.v_out + (int64_t)bh * C * RB;
        for (int t = warp; t < 2 * nbc; t += nsw) {
            const bool isv = t >= nbc;
            const int j0 = (isv ? t - nbc : t) << 5;
            const int rows = min(32, C - j0);
            const int j = j0 + lane;
            int row = 0;
            if (lane < rows) row = j < sink ? j : (j < sink + ksel ? sidx[j - sink] : j + tail0);
            const char* src = isv ? vbase + (int64_t)row * L.vss : kbase + (int64_t)row * L.kss;
            warp_load_rows<RB>(slot, bar, src, rows, isv ? vdense : kdense, lane);
            mbar_wait(bar, parity);
            parity ^= 1;
            if (lane == 0) {
                bulk_s2g((isv ? vo : ko) + (int64_t)j0 * RB, slot, (uint32_t)rows * RB);
                bulk_commit();
                bulk_wait_read<0>();  // the slot has been read out: it may be refilled
            }
            __syncwarp();
        }
    }
    __syncthreads();  // every warp is done with sidx / keys before the next unit reuses them
    }  // unit loop
}

}  // namespace kvc
