// kvc_vote_split.cuh — persistent, split-sequence form of the TMA-fed tcgen05 vote kernel (kvc_vote.cuh).
//
// Why: the vote needs two passes over the keys of a (batch, KV head) "unit" — pass 1 for the softmax row statistics,
// pass 2 for the per-key votes — and a 32K-token unit is 8 MB of keys.  With one CTA per unit and 148 units in
// flight the second read never hits L2: the kernel reads K twice from HBM (profiles/r01_vote_ncu_full_c4_b4.json:
// 17.2 GB for 8.6 GB of keys) and sits at ~80 % of the measured copy bandwidth.  Here a unit is cut along S into
// slices of a few hundred KB, each slice is read by ONE CTA for both passes, and the slices of a unit exchange
// their row statistics through a small device workspace — the second read comes out of L2.
//
// Shape of the kernel: one persistent CTA per SM, six roles.
//   warp 16 lane 0   producer   draws slice tickets, decides the phase order, issues every TMA load (Q and K tiles)
//   warp 17 lane 0   MMA        tcgen05.mma into four 128-column TMEM accumulators
//   warps 0-15       math       four groups, one accumulator each: softmax statistics (pass 1) / votes (pass 2)
//   warp 18          publisher  merges the four groups' rows of a finished pass-1 phase, publishes them, counts the
//                               slice in; the LAST slice of a unit to arrive merges all slices (in slice order, so the
//                               result does not depend on arrival order) and raises the unit's ready flag
//   warp 19          prefetch   for every pass-2 phase waits for the ready flag and stages the unit's final row
//                               statistics in shared memory before the math groups get there
// The producer keeps up to pend_max slices between their passes and queues a pass 2 as soon as its unit is ready, or
// -- ready or not -- when the pending list is full or the tickets have run out (the tiles are staged anyway; the math
// warps pick the statistics up when they get there).  Forward progress: tickets are drawn in start order, and no
// ticket is drawn while an unready pass 2 sits in the queue, so the pass 1 of every ticket a CTA holds is queued
// AHEAD of anything that can wait; the CTAs that wait all hold tickets of the one unit that straddles the ticket
// counter, and with ns <= 64 slices per unit (host cap) some CTA is always free to draw the rest as long as
// 64 / pend_max CTAs are resident.  Every wait is bounded and traps instead of hanging.
//
// STATUS (round 1): correct (tests/test_gpu_vote.py) and the HBM traffic does halve -- 8.6 GB instead of 17.2 GB on
// the B = 4 profile -- but only for slices of <= 6 tiles (the L2 holds ~60 MB of single-reader lines, ~12 tiles per
// SM), and at that size every phase pays ~3 us of cross-CTA latency (publish -> merge -> flag -> prefetch) that a
// 4-6 tile phase cannot hide: 16.3-18.5 ms against 15.4 ms for one CTA per unit on c4_vote.  Shipped OPT-IN
// (KVC_VOTE_SPLIT=1); profiles/r01_vote_split_sweep.json has the sweep.
#pragma once
#include "kvc_vote.cuh"

namespace kvc {

constexpr int kSpRing = 4;      // key-tile ring slots (slot == accumulator index)
constexpr int kSpDesc = 32;     // phase descriptor queue (the producer is < 16 phases ahead of the slowest reader)
constexpr int kSpPend = 3;      // slices a CTA may hold between pass 1 and pass 2
constexpr int kSpThreads = 640;
constexpr int kSpHeader = 12288;  // barriers, descriptors, statistics; the operand buffers follow (1024-aligned)

struct VoteSplitLayerDev {
    alignas(64) CUtensorMap kmap;       // keys [B,H,S,D] as (D, S, H, B), box (64, 128, 1, 1), SWIZZLE_128B
    alignas(64) CUtensorMap kmap_tail;  // D % 64 == 16: box (16, 128, 1, 1), SWIZZLE_32B
    alignas(64) CUtensorMap qmap;       // queries [B,H*G,W,D] as (D, W, H*G, B), box (64, W, G, 1), SWIZZLE_128B
    alignas(64) CUtensorMap qmap_tail;  // box (16, W, G, 1), SWIZZLE_32B
    char* votes;
    int32_t S, first, ns, pad;  // first = ticket of this layer's first slice; ns = slices per unit
};
struct VoteSplitBatchDev {
    int32_t B, H, G, W;
    float scale_log2e;
    int32_t n_layers, total, pend_max;  // pend_max: slices a CTA may hold between their passes (1..kSpPend)
    int32_t per_layer, pad[3];          // tickets per layer when all layers agree, else 0
    uint32_t* ws;                 // [0] ticket counter, [16 + unit] slices arrived, [16 + units + unit] ready flag
    int64_t off_final, off_part;  // byte offsets into ws: float final[unit][2][128], float part[ticket][2][128]
    VoteSplitLayerDev layers[32];
};

struct VoteItem {
    int layer, bh, unit, slice, ns, t_begin, n1, n2, S, P;
};
// What the producer tells the other warps about a phase (one pass over one slice): 32 bytes in shared memory.
struct alignas(16) VotePhase {
    int ticket;       // -1: no more phases
    int pass;         // 0: statistics (tiles t_begin + i), 1: votes (tiles t_begin + nt - 1 - i)
    int layer, bh;
    int t_begin, nt;
    int unit, slice;
};
// ticket -> (layer, unit, slice) and the slice's tile range; tiles are spread evenly over the slices of a unit
__device__ __forceinline__ VoteItem vote_item(const VoteSplitBatchDev& bd, int ticket) {
    int layer = 0;
    if (bd.per_layer > 0)
        layer = ticket / bd.per_layer;  // every layer has the same number of slices (the usual case)
    else
        while (layer + 1 < bd.n_layers && ticket >= bd.layers[layer + 1].first) ++layer;
    const VoteSplitLayerDev& L = bd.layers[layer];
    VoteItem it;
    it.layer = layer;
    it.S = L.S;
    it.P = L.S - bd.W;
    it.ns = L.ns;
    const int rem = ticket - L.first;
    it.bh = rem / it.ns;
    it.slice = rem - it.bh * it.ns;
    it.unit = layer * (bd.B * bd.H) + it.bh;
    const int n1_all = (it.S + kVoteTile - 1) / kVoteTile, n2_all = (it.P + kVoteTile - 1) / kVoteTile;
    const int base = n1_all / it.ns, extra = n1_all - base * it.ns;
    it.t_begin = it.slice * base + min(it.slice, extra);
    it.n1 = base + (it.slice < extra ? 1 : 0);       // pass-1 tiles: t_begin + i
    it.n2 = max(0, min(n2_all - it.t_begin, it.n1));  // pass-2 tiles (keys before the window), walked backwards
    return it;
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int DT, int CPR>
__global__ void __launch_bounds__(kSpThreads, 1) kvc_snapkv_vote_split_kernel(const __grid_constant__ VoteSplitBatchDev bd) {
    using Tr = Traits<DT>;
    using Key = typename Tr::Key;
    static_assert(DT != KVC_DTYPE_F32, "the vote runs on 16-bit caches (kind::f16)");
    static_assert(CPR % 8 == 0 || CPR % 8 == 2, "rows = 128-byte boxes (+ one 32-byte box for D % 64 == 16)");
    constexpr int KH = CPR / 8;
    constexpr int REM = CPR % 8;
    constexpr int BOX_BYTES = kVoteTile * 128;
    constexpr int TAIL_BYTES = kVoteTile * 16 * REM;
    constexpr int TILE_BYTES = KH * BOX_BYTES + TAIL_BYTES;  // a key tile and the Q operand have the same layout
    constexpr uint32_t IDESC = umma_idesc_f16(DT == KVC_DTYPE_BF16 ? 1 : 0, kVoteM, kVoteTile);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = bd.W, G = bd.G;
    const int rows_q = G * W;
    const int units = bd.n_layers * bd.B * bd.H;

    extern __shared__ __align__(1024) unsigned char smem_split[];
    unsigned char* smem = smem_split;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem);
    const uint32_t bar_full = smem_u32(smem + 64);          // [ring]  1 + TMA transaction bytes
    const uint32_t bar_empty = bar_full + 8 * kSpRing;      // [ring]  1 (tcgen05.commit)
    const uint32_t bar_tfull = bar_empty + 8 * kSpRing;     // [4]     1 (tcgen05.commit)
    const uint32_t bar_tempty = bar_tfull + 8 * 4;          // [4]     128 (one math group)
    const uint32_t bar_qfull = bar_tempty + 8 * 4;          // [2]     1 + TMA transaction bytes
    const uint32_t bar_qempty = bar_qfull + 8 * 2;          // [2]     1 (tcgen05.commit)
    const uint32_t bar_sfull = bar_qempty + 8 * 2;          // [2]     32 (prefetch warp)
    const uint32_t bar_sempty = bar_sfull + 8 * 2;          // [2]     512 (math)
    const uint32_t bar_pfull = bar_sempty + 8 * 2;          // [2]     512 (math)
    const uint32_t bar_pempty = bar_pfull + 8 * 2;          // [2]     32 (publisher warp)
    const uint32_t bar_desc = bar_pempty + 8 * 2;           // [kSpDesc] 1 (producer)
    static_assert(64 + 8 * (2 * kSpRing + 8 + 12 + kSpDesc) <= 1024, "barriers overflow their slab");
    VotePhase* s_desc = reinterpret_cast<VotePhase*>(smem + 1024);          // [kSpDesc] phase descriptors
    static_assert(sizeof(VotePhase) * kSpDesc <= 1024, "descriptor queue overflows its slab");
    float* s_stat = reinterpret_cast<float*>(smem + 2048);                  // [2][2][128] final (m, 1/l) of a unit
    float* s_part = reinterpret_cast<float*>(smem + 4096);                  // [2][4 groups][2][128] pass-1 rows
    unsigned char* s_q = smem + kSpHeader;                                  // [2] Q operands
    unsigned char* s_ring = s_q + 2 * TILE_BYTES;

    if (tid == 0) {
        for (int i = 0; i < kSpRing; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 128);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_qfull + 8 * i, 1);
            mbar_init(bar_qempty + 8 * i, 1);
            mbar_init(bar_sfull + 8 * i, 32);
            mbar_init(bar_sempty + 8 * i, 512);
            mbar_init(bar_pfull + 8 * i, 512);
            mbar_init(bar_pempty + 8 * i, 32);
        }
        for (int i = 0; i < kSpDesc; ++i) mbar_init(bar_desc + 8 * i, 1);
        mbar_init_fence();
    }
    if (warp == 17) tmem_alloc(smem_u32(s_tmem), 512);
    // query rows beyond G*W are never written by the TMA loads: zero them once
    for (int i = tid; i < 2 * TILE_BYTES / 16; i += kSpThreads) reinterpret_cast<int4*>(s_q)[i] = make_int4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t q_addr = smem_u32(s_q), ring_addr = smem_u32(s_ring);
    const float c2 = bd.scale_log2e;
    uint32_t* ws_ready = bd.ws + 16 + units;
    float* ws_final = reinterpret_cast<float*>(reinterpret_cast<char*>(bd.ws) + bd.off_final);
    float* ws_part = reinterpret_cast<float*>(reinterpret_cast<char*>(bd.ws) + bd.off_part);

    if (warp < 16) {
        // ================================================================ math groups
        const int grp = warp >> 2, gt = tid & 127;
        const uint32_t t_lane = tmem + grp * kVoteTile + ((uint32_t)((warp & 3) * 32) << 16);
        const bool row_live = (warp & 3) * 32 < rows_q;
        const int w_of_row = gt < rows_q ? gt % W : 0;
        uint32_t va[16], vb[16];  // the accumulator is read 16 columns at a time, the next 16 already in flight
        uint32_t g = 0;
        int np1 = 0, np2 = 0;
        for (int p = 0;; ++p) {
            mbar_wait(bar_desc + 8 * (p & (kSpDesc - 1)), (uint32_t)((p / kSpDesc) & 1));
            const VotePhase ph = s_desc[p & (kSpDesc - 1)];
            if (ph.ticket < 0) break;
            const int S = bd.layers[ph.layer].S, P = S - W;
            const int first_i = (int)((grp - g) & 3u);
            if (ph.pass == 0) {
                // ---------------- pass 1: lane = query row, columns = keys of a tile (online softmax statistics)
                float m_run = -INFINITY, l_run = 0.f;
                const int limit = P + w_of_row;
                for (int i = first_i; i < ph.nt; i += 4) {
                    mbar_wait(bar_tfull + 8 * grp, (uint32_t)(((g + i) >> 2) & 1));
                    tc_fence_after();
                    if (row_live) {
                        const int key0 = (ph.t_begin + i) * kVoteTile;
                        const bool masked = key0 + kVoteTile > P;
                        tmem_ld16_async(t_lane, va);
#pragma unroll
                        for (int cb = 0; cb < kVoteTile; cb += 16) {
                            uint32_t(&v)[16] = ((cb >> 4) & 1) ? vb : va;
                            tmem_ld_wait();
                            if (cb + 16 < kVoteTile) tmem_ld16_async(t_lane + cb + 16, ((cb >> 4) & 1) ? va : vb);
                            float cmax = -INFINITY;
                            if (masked) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const int key = key0 + cb + j;
                                    if (key > limit || key >= S) v[j] = 0xff800000u;
                                    cmax = fmaxf(cmax, __uint_as_float(v[j]));
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) cmax = fmaxf(cmax, __uint_as_float(v[j]));
                            }
                            const float m_new = fmaxf(m_run, cmax * c2);
                            if (m_new > -INFINITY) {
                                float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    acc0 += ex2(fmaf(__uint_as_float(v[j]), c2, -m_new));
                                    acc1 += ex2(fmaf(__uint_as_float(v[j + 1]), c2, -m_new));
                                    acc2 += ex2(fmaf(__uint_as_float(v[j + 2]), c2, -m_new));
                                    acc3 += ex2(fmaf(__uint_as_float(v[j + 3]), c2, -m_new));
                                }
                                if (m_new != m_run) l_run *= ex2(m_run - m_new);  // the maximum settles quickly
                                l_run += (acc0 + acc1) + (acc2 + acc3);
                                m_run = m_new;
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(bar_tempty + 8 * grp);
                }
                g += (uint32_t)ph.nt;
                // hand this group's rows to the publisher warp
                const int pb = np1 & 1;
                mbar_wait(bar_pempty + 8 * pb, (uint32_t)(((np1 >> 1) & 1) ^ 1));  // fresh barrier: passes
                float* sp = s_part + pb * 1024 + grp * 256;
                sp[gt] = m_run;
                sp[128 + gt] = l_run;
                mbar_arrive(bar_pfull + 8 * pb);
                ++np1;
            } else {
                // ---------------- pass 2: lane = key, columns = query rows
                const int sb = np2 & 1;
                mbar_wait(bar_sfull + 8 * sb, (uint32_t)((np2 >> 1) & 1));
                const float* s_m = s_stat + sb * 256;
                const float* s_invl = s_m + 128;
                Key* out = reinterpret_cast<Key*>(bd.layers[ph.layer].votes) + (int64_t)ph.bh * P;
                for (int i = first_i; i < ph.nt; i += 4) {
                    mbar_wait(bar_tfull + 8 * grp, (uint32_t)(((g + i) >> 2) & 1));
                    tc_fence_after();
                    float vote0 = 0.f, vote1 = 0.f, vote2 = 0.f, vote3 = 0.f;
                    tmem_ld16_async(t_lane, va);
#pragma unroll
                    for (int cb = 0; cb < kVoteM; cb += 16) {
                        if (cb < rows_q) {  // warp-uniform: padding query rows never vote
                            uint32_t(&v)[16] = ((cb >> 4) & 1) ? vb : va;
                            tmem_ld_wait();
                            if (cb + 16 < rows_q) tmem_ld16_async(t_lane + cb + 16, ((cb >> 4) & 1) ? va : vb);
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 mm = *reinterpret_cast<const float4*>(s_m + cb + j);
                                const float4 il = *reinterpret_cast<const float4*>(s_invl + cb + j);
                                vote0 = fmaf(ex2(fmaf(__uint_as_float(v[j + 0]), c2, -mm.x)), il.x, vote0);
                                vote1 = fmaf(ex2(fmaf(__uint_as_float(v[j + 1]), c2, -mm.y)), il.y, vote1);
                                vote2 = fmaf(ex2(fmaf(__uint_as_float(v[j + 2]), c2, -mm.z)), il.z, vote2);
                                vote3 = fmaf(ex2(fmaf(__uint_as_float(v[j + 3]), c2, -mm.w)), il.w, vote3);
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(bar_tempty + 8 * grp);
                    const int key = (ph.t_begin + ph.nt - 1 - i) * kVoteTile + gt;
                    if (key < P) out[key] = (Key)Tr::to_raw((vote0 + vote1) + (vote2 + vote3));
                }
                g += (uint32_t)ph.nt;
                mbar_arrive(bar_sempty + 8 * sb);
                ++np2;
            }
        }
    } else if (warp == 16) {
        // ================================================================ producer (one thread)
        if (lane == 0) {
            uint32_t g = 0;
            int p = 0;
            int pend0 = -1, pend1 = -1, pend2 = -1, npend = 0;  // slices between their passes, oldest first
            int unit0 = 0, unit1 = 0, unit2 = 0;                // ... and their units (ready-flag index)
            // Global round trips (ticket counter, ready flag) are issued at the START of a phase's tile loop and
            // consumed at the next decision, one phase later: the producer never sits on a 1 us atomic between
            // two phases while the ring runs dry.
            int next_t = (int)atomicAdd(bd.ws, 1u);  // pre-drawn ticket, -1: none held
            bool exhausted = false;
            uint32_t head_flag = 0;                  // last polled value of ready[unit0]
            int guard = -1;                          // unit of a pass 2 that was queued before its unit was ready
            if (next_t >= bd.total) next_t = -1, exhausted = true;
            // No ticket is drawn while `guard` is unready: a new ticket could be a sibling of the slice whose pass 2
            // is waiting in the queue, and its pass 1 would sit behind that pass 2 forever.
            auto guard_clear = [&](bool block) -> bool {
                if (guard < 0) return true;
                const uint32_t* flag = ws_ready + guard;
                bool ok = ld_acquire_gpu(flag) != 0;
                if (!ok && block) {
                    for (int spin = 0; spin < (1 << 25) && !ok; ++spin) {
                        __nanosleep(100);
                        ok = ld_acquire_gpu(flag) != 0;
                    }
                    if (!ok) __trap();  // a sibling slice never arrived: fail the launch instead of hanging
                }
                if (ok) guard = -1;
                return ok;
            };
            for (;;) {
                int ticket = -1, pass = 0;
                if (npend > 0) {
                    // Pass 2 of the oldest pending slice is queued as soon as its unit is ready, or -- ready or not --
                    // when the pending list is full or the tickets have run out.  The producer runs ~8 tiles ahead of
                    // the math warps, so "not ready yet" is the common case: the tiles are staged anyway and the math
                    // warps pick the statistics up when they get there.  A pass 2 is only queued unready when no
                    // un-queued ticket is held (tickets are pre-drawn only while the list has room).
                    const bool ready = head_flag != 0;
                    if (ready || npend >= bd.pend_max || (next_t < 0 && exhausted)) {
                        if (!ready) {
                            guard_clear(true);  // at most one unready pass 2 in the queue
                            guard = unit0;
                        }
                        ticket = pend0, pass = 1;
                        pend0 = pend1, pend1 = pend2, pend2 = -1;
                        unit0 = unit1, unit1 = unit2;
                        head_flag = 0;
                        --npend;
                    }
                }
                if (ticket < 0) {
                    if (next_t < 0) {
                        if (exhausted) break;  // nothing pending (a pending slice would have been taken), nothing left
                        guard_clear(true);
                        next_t = (int)atomicAdd(bd.ws, 1u);  // not pre-drawn: draw now
                        if (next_t >= bd.total) {
                            next_t = -1, exhausted = true;
                            continue;
                        }
                    }
                    ticket = next_t;
                    next_t = -1;
                }
                const VoteItem it = vote_item(bd, ticket);
                if (pass == 0) {
                    if (npend == 0) pend0 = ticket, unit0 = it.unit;
                    else if (npend == 1) pend1 = ticket, unit1 = it.unit;
                    else pend2 = ticket, unit2 = it.unit;
                    ++npend;
                }
                const int nt = pass ? it.n2 : it.n1;
                if (nt == 0) continue;  // a slice of window keys only has no pass 2
                const VoteSplitLayerDev& L = bd.layers[it.layer];
                const int b = it.bh / bd.H, h = it.bh - b * bd.H;
                VotePhase ph;
                ph.ticket = ticket, ph.pass = pass, ph.layer = it.layer, ph.bh = it.bh;
                ph.t_begin = it.t_begin, ph.nt = nt, ph.unit = it.unit, ph.slice = it.slice;
                s_desc[p & (kSpDesc - 1)] = ph;
                mbar_arrive(bar_desc + 8 * (p & (kSpDesc - 1)));
                // this phase's Q operand
                const int qs = p & 1;
                mbar_wait(bar_qempty + 8 * qs, (uint32_t)(((p >> 1) & 1) ^ 1));  // fresh barrier: passes
                mbar_arrive_expect_tx(bar_qfull + 8 * qs, (uint32_t)(rows_q * (KH * 128 + REM * 16)));
#pragma unroll
                for (int kh = 0; kh < KH; ++kh)
                    tma_load_4d(q_addr + qs * TILE_BYTES + kh * BOX_BYTES, &L.qmap, kh * 64, 0, h * G, b, bar_qfull + 8 * qs);
                if (REM > 0)
                    tma_load_4d(q_addr + qs * TILE_BYTES + KH * BOX_BYTES, &L.qmap_tail, KH * 64, 0, h * G, b,
                                bar_qfull + 8 * qs);
                // next decision's inputs, in flight while this phase's tiles are issued
                uint32_t drawn = 0xffffffffu, polled = 0;
                const bool draw = next_t < 0 && !exhausted && npend < bd.pend_max && guard_clear(false);
                if (draw) drawn = atomicAdd(bd.ws, 1u);
                if (npend > 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(polled) : "l"(ws_ready + unit0));
                for (int i = 0; i < nt; ++i, ++g) {
                    const int slot = (int)(g % kSpRing);
                    mbar_wait(bar_empty + 8 * slot, (uint32_t)(((g / kSpRing) & 1) ^ 1));  // fresh barrier: passes
                    const int t = it.t_begin + (pass ? nt - 1 - i : i);
                    mbar_arrive_expect_tx(bar_full + 8 * slot, TILE_BYTES);
#pragma unroll
                    for (int kh = 0; kh < KH; ++kh)  // rows beyond S are zero-filled by the TMA unit
                        tma_load_4d(ring_addr + slot * TILE_BYTES + kh * BOX_BYTES, &L.kmap, kh * 64, t * kVoteTile, h, b,
                                    bar_full + 8 * slot);
                    if (REM > 0)
                        tma_load_4d(ring_addr + slot * TILE_BYTES + KH * BOX_BYTES, &L.kmap_tail, KH * 64, t * kVoteTile, h,
                                    b, bar_full + 8 * slot);
                }
                if (draw) {
                    if ((int)drawn >= bd.total) exhausted = true;
                    else next_t = (int)drawn;
                }
                head_flag = npend > 0 ? polled : 0;
                ++p;
            }
            VotePhase end;
            end.ticket = -1, end.pass = 0, end.layer = 0, end.bh = 0, end.t_begin = 0, end.nt = 0, end.unit = 0, end.slice = 0;
            s_desc[p & (kSpDesc - 1)] = end;
            mbar_arrive(bar_desc + 8 * (p & (kSpDesc - 1)));
        }
    } else if (warp == 17) {
        // ================================================================ MMA issuer (one thread)
        if (lane == 0) {
            uint32_t g = 0;
            for (int p = 0;; ++p) {
                mbar_wait(bar_desc + 8 * (p & (kSpDesc - 1)), (uint32_t)((p / kSpDesc) & 1));
                const VotePhase ph = s_desc[p & (kSpDesc - 1)];
                if (ph.ticket < 0) break;
                const int nt = ph.nt, pass = ph.pass;
                const int qs = p & 1;
                mbar_wait(bar_qfull + 8 * qs, (uint32_t)((p >> 1) & 1));
                const uint32_t qb = q_addr + qs * TILE_BYTES;
                for (int i = 0; i < nt; ++i, ++g) {
                    const int slot = (int)(g % kSpRing), acc = (int)(g & 3);
                    mbar_wait(bar_tempty + 8 * acc, (uint32_t)(((g >> 2) & 1) ^ 1));  // accumulator drained
                    mbar_wait(bar_full + 8 * slot, (uint32_t)((g / kSpRing) & 1));    // tile landed
                    tc_fence_after();
                    const uint32_t kb = ring_addr + slot * TILE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < CPR / 2; ++ks) {
                        // both operands: 128B-swizzled K-major boxes (8-row groups 1024 B apart), 32 bytes per K step
                        const bool body = (ks >> 2) < KH;
                        const uint32_t off = body ? (ks >> 2) * BOX_BYTES + (ks & 3) * 32 : KH * BOX_BYTES;
                        const uint64_t kd = body ? umma_smem_desc_sw128(kb + off) : umma_smem_desc_sw32(kb + off);
                        const uint64_t qd = body ? umma_smem_desc_sw128(qb + off) : umma_smem_desc_sw32(qb + off);
                        umma_f16(tmem + acc * kVoteTile, pass ? kd : qd, pass ? qd : kd, IDESC, ks > 0 ? 1u : 0u);
                    }
                    umma_commit(bar_tfull + 8 * acc);
                    umma_commit(bar_empty + 8 * slot);
                }
                umma_commit(bar_qempty + 8 * qs);
            }
        }
    } else if (warp == 18) {
        // ================================================================ publisher: pass-1 rows -> workspace
        int np1 = 0;
        for (int p = 0;; ++p) {
            mbar_wait(bar_desc + 8 * (p & (kSpDesc - 1)), (uint32_t)((p / kSpDesc) & 1));
            const VotePhase ph = s_desc[p & (kSpDesc - 1)];
            if (ph.ticket < 0) break;
            if (ph.pass != 0) continue;
            const int ticket = ph.ticket, ns = bd.layers[ph.layer].ns;
            const int pb = np1 & 1;
            mbar_wait(bar_pfull + 8 * pb, (uint32_t)((np1 >> 1) & 1));
            const float* sp = s_part + pb * 1024;
            float mg[4][4], lg[4][4];  // [group][row of this lane]
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
                const float4 a = *reinterpret_cast<const float4*>(sp + g4 * 256 + 4 * lane);
                const float4 c = *reinterpret_cast<const float4*>(sp + g4 * 256 + 128 + 4 * lane);
                mg[g4][0] = a.x, mg[g4][1] = a.y, mg[g4][2] = a.z, mg[g4][3] = a.w;
                lg[g4][0] = c.x, lg[g4][1] = c.y, lg[g4][2] = c.z, lg[g4][3] = c.w;
            }
            mbar_arrive(bar_pempty + 8 * pb);
            ++np1;
            float m[4], l[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                m[r] = fmaxf(fmaxf(mg[0][r], mg[1][r]), fmaxf(mg[2][r], mg[3][r]));
                l[r] = 0.f;
                if (m[r] > -INFINITY) {
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4)
                        if (mg[g4][r] > -INFINITY) l[r] += lg[g4][r] * ex2(mg[g4][r] - m[r]);
                }
            }
            float* mine = ws_part + (int64_t)ticket * 256;
            __stcg(reinterpret_cast<float4*>(mine + 4 * lane), make_float4(m[0], m[1], m[2], m[3]));
            __stcg(reinterpret_cast<float4*>(mine + 128 + 4 * lane), make_float4(l[0], l[1], l[2], l[3]));
            __threadfence();
            __syncwarp();
            uint32_t old = 0;
            if (lane == 0) old = atomicAdd(bd.ws + 16 + ph.unit, 1u);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == (uint32_t)(ns - 1)) {
                // last slice of the unit to arrive: merge every slice's rows, in slice order
                __threadfence();
                const float* sib = ws_part + (int64_t)(ticket - ph.slice) * 256;
                float M[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, Lsum[4] = {0.f, 0.f, 0.f, 0.f};
                for (int j = 0; j < ns; ++j) {
                    const float4 a = __ldcg(reinterpret_cast<const float4*>(sib + (int64_t)j * 256 + 4 * lane));
                    M[0] = fmaxf(M[0], a.x), M[1] = fmaxf(M[1], a.y), M[2] = fmaxf(M[2], a.z), M[3] = fmaxf(M[3], a.w);
                }
                for (int j = 0; j < ns; ++j) {
                    const float4 a = __ldcg(reinterpret_cast<const float4*>(sib + (int64_t)j * 256 + 4 * lane));
                    const float4 c = __ldcg(reinterpret_cast<const float4*>(sib + (int64_t)j * 256 + 128 + 4 * lane));
                    if (a.x > -INFINITY) Lsum[0] += c.x * ex2(a.x - M[0]);
                    if (a.y > -INFINITY) Lsum[1] += c.y * ex2(a.y - M[1]);
                    if (a.z > -INFINITY) Lsum[2] += c.z * ex2(a.z - M[2]);
                    if (a.w > -INFINITY) Lsum[3] += c.w * ex2(a.w - M[3]);
                }
                float fm[4], fi[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const bool live = 4 * lane + r < rows_q && Lsum[r] > 0.f;
                    fm[r] = live ? M[r] : 0.f;
                    fi[r] = live ? 1.f / Lsum[r] : 0.f;
                }
                float* f = ws_final + (int64_t)ph.unit * 256;
                __stcg(reinterpret_cast<float4*>(f + 4 * lane), make_float4(fm[0], fm[1], fm[2], fm[3]));
                __stcg(reinterpret_cast<float4*>(f + 128 + 4 * lane), make_float4(fi[0], fi[1], fi[2], fi[3]));
                __threadfence();
                __syncwarp();
                if (lane == 0) st_release_gpu(ws_ready + ph.unit, 1u);
            }
        }
    } else if (warp == 19) {
        // ================================================================ prefetch: final rows of a unit -> smem
        int np2 = 0;
        for (int p = 0;; ++p) {
            mbar_wait(bar_desc + 8 * (p & (kSpDesc - 1)), (uint32_t)((p / kSpDesc) & 1));
            const VotePhase ph = s_desc[p & (kSpDesc - 1)];
            if (ph.ticket < 0) break;
            if (ph.pass == 0) continue;
            if (lane == 0) {
                uint32_t seen = 0;
                for (int spin = 0; spin < (1 << 25); ++spin) {
                    seen = ld_acquire_gpu(ws_ready + ph.unit);
                    if (seen) break;
                    __nanosleep(100);
                }
                if (!seen) __trap();  // a sibling slice never arrived: fail the launch instead of hanging
            }
            __syncwarp();
            __threadfence();
            const float* f = ws_final + (int64_t)ph.unit * 256;
            const float4 a = __ldcg(reinterpret_cast<const float4*>(f + 4 * lane));
            const float4 c = __ldcg(reinterpret_cast<const float4*>(f + 128 + 4 * lane));
            const int sb = np2 & 1;
            mbar_wait(bar_sempty + 8 * sb, (uint32_t)(((np2 >> 1) & 1) ^ 1));  // fresh barrier: passes
            *reinterpret_cast<float4*>(s_stat + sb * 256 + 4 * lane) = a;
            *reinterpret_cast<float4*>(s_stat + sb * 256 + 128 + 4 * lane) = c;
            mbar_arrive(bar_sfull + 8 * sb);
            ++np2;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem, 512);
}

}  // namespace kvc
