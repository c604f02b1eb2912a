/*
 * kvc_oracle.c — plain-C restatement of the reference's per-layer compress step.
 * TEST INFRASTRUCTURE, NOT PRODUCT: used by tests/ as a second checker (cross-validated against
 * oracle/kvc_oracle.py, which is pinned to golden vectors of the real reference) and by bench.py
 * as the CPU baseline ("port").  Nothing under cs3602-llm-inference-acceleration_b200/ links or
 * loads this file.
 *
 * It follows the reference's own steps on purpose (it stands in for the reference's CPU cost, it
 * is not an optimised algorithm):
 *     norm over the selection region          torch.norm(K, p=2, dim=-1)      l2_compress.py:70
 *     FULL sort of the region by key          token_norms.argsort(dim=-1)     l2_compress.py:73
 *     take the first k                        sorted_indices[:, :, :k]        l2_compress.py:76
 *     sort the kept indices ascending         torch.sort(indices_to_keep)     l2_compress.py:79
 *     gather K and V rows                     torch.gather x2                 l2_compress.py:87-88
 *     sinks + selected + recent               torch.cat x2                    h2o_l2.py:148-149
 * (the same skeleton in fix_size_l2.py:104-147, h2o_l2.py:122-149, pyramid_kv.py:155-181,
 *  adaptive_l2.py:126-143,180-197; snapkv_lite.py:96-150 with the score pipeline below).
 * The reference runs these as multi-threaded ATen CPU kernels; here OpenMP parallelises over
 * (batch, head) rows.  Ties: lowest token index first (stable order) — see kvc_oracle.py.
 *
 * Build: gcc -O3 -fopenmp -fPIC -shared -o libkvc_oracle.so kvc_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };
enum { MODE_NONE = 0, MODE_LOW = 1, MODE_HIGH = 2, MODE_SNAPKV = 3 };

static inline float bits_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline float bf16_to_f32(uint16_t h) { return bits_f32((uint32_t)h << 16); }
static inline uint16_t f32_to_bf16(float f) { /* round to nearest even */
    uint32_t u = f32_bits(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
static inline float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
    if (exp == 0) {
        if (man == 0) return bits_f32(sign);
        float v = (float)man * 5.9604644775390625e-8f; /* 2^-24 */
        return sign ? -v : v;
    }
    if (exp == 31) return bits_f32(sign | 0x7f800000u | (man << 13));
    return bits_f32(sign | ((exp + 112u) << 23) | (man << 13));
}
static inline uint16_t f32_to_f16(float f) { /* round to nearest even */
    uint32_t u = f32_bits(f), sign = (u >> 16) & 0x8000u, a = u & 0x7fffffffu;
    if (a > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);
    if (a >= 0x47800000u) return (uint16_t)(sign | 0x7c00u); /* >= 65536 -> inf (65520 rounds to inf below) */
    if (a < 0x38800000u) { /* subnormal half or zero */
        if (a < 0x33000000u) return (uint16_t)sign; /* < 2^-25 */
        uint32_t e = a >> 23, m = (a & 0x7fffffu) | 0x800000u;
        uint32_t shift = 126u - e; /* 14..24 */
        uint32_t r = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1u))) r++;
        return (uint16_t)(sign | r);
    }
    uint32_t r = ((a - 0x38000000u) >> 13), rem = a & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;
    return (uint16_t)(sign | r); /* carry into the exponent (up to inf) is the correct rounding */
}

static inline float load_elem(const void* p, int dtype, int64_t i) {
    if (dtype == DT_F32) return ((const float*)p)[i];
    if (dtype == DT_F16) return f16_to_f32(((const uint16_t*)p)[i]);
    return bf16_to_f32(((const uint16_t*)p)[i]);
}
static inline float round_dt(float x, int dtype) {
    if (dtype == DT_F32) return x;
    if (dtype == DT_F16) return f16_to_f32(f32_to_f16(x));
    return bf16_to_f32(f32_to_bf16(x));
}
static inline int elem_size(int dtype) { return dtype == DT_F32 ? 4 : 2; }

typedef struct { float key; int32_t idx; } keyed;

static int cmp_keyed(const void* a, const void* b) {
    const keyed* x = (const keyed*)a; const keyed* y = (const keyed*)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx); /* ties: lowest index first */
}
static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

/* norms of rows [lo, hi) of one (b,h) slab: fp32 accumulation, rounded once to the dtype */
static void row_norms(const void* k, int dtype, int D, int lo, int hi, float* out) {
    for (int r = lo; r < hi; ++r) {
        float acc = 0.f;
        const int64_t base = (int64_t)r * D;
        for (int d = 0; d < D; ++d) { float x = load_elem(k, dtype, base + d); acc += x * x; }
        out[r - lo] = round_dt(sqrtf(acc), dtype);
    }
}

/* snapkv_lite.py:96-121: (max + 1e-6) - norm in the dtype, avg_pool1d(kernel, 1, kernel/2) */
static void snapkv_scores(float* s, float* tmp, int n, int dtype, int kernel) {
    float mx = s[0];
    for (int i = 1; i < n; ++i) if (s[i] > mx) mx = s[i];
    const float mxe = round_dt(mx + 1e-6f, dtype);
    for (int i = 0; i < n; ++i) s[i] = round_dt(mxe - s[i], dtype);
    if (kernel > 1 && n >= kernel) {
        const int pad = kernel / 2;
        for (int i = 0; i < n; ++i) {
            float acc = 0.f;
            for (int t = 0; t < kernel; ++t) { int j = i - pad + t; if (j >= 0 && j < n) acc += s[j]; }
            tmp[i] = round_dt(acc / (float)kernel, dtype);
        }
        memcpy(s, tmp, (size_t)n * sizeof(float));
    }
}

/*
 * One layer. k_in/v_in: dense [B,H,S,D]; k_out/v_out: dense [B,H,C,D], C = sink + k_sel + tail;
 * rows_out: optional [B,H,C] kept rows. Returns 0, or -1 on bad arguments / allocation failure.
 */
int kvc_oracle_layer(int dtype, int B, int H, int S, int D, const void* k_in, const void* v_in, int sink, int lo,
                     int hi, int k_sel, int tail, int mode, int pool_kernel, void* k_out, void* v_out,
                     int32_t* rows_out, int nthreads) {
    if (dtype < 0 || dtype > 2 || B <= 0 || H <= 0 || S < 0 || D <= 0) return -1;
    if (sink < 0 || tail < 0 || k_sel < 0 || sink > S || tail > S) return -1;
    if (k_sel > 0 && (lo < 0 || hi > S || hi - lo < k_sel || mode == MODE_NONE)) return -1;
    const int C = sink + k_sel + tail, R = k_sel > 0 ? hi - lo : 0;
    const int es = elem_size(dtype);
    const size_t row_bytes = (size_t)D * es;
    int failed = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        float* score = (float*)malloc(sizeof(float) * (size_t)(R > 0 ? R : 1) * 2);
        keyed* order = (keyed*)malloc(sizeof(keyed) * (size_t)(R > 0 ? R : 1));
        int32_t* rows = (int32_t*)malloc(sizeof(int32_t) * (size_t)(C > 0 ? C : 1));
        if (!score || !order || !rows) {
#pragma omp atomic write
            failed = 1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int bh = 0; bh < B * H; ++bh) {
                const char* kb = (const char*)k_in + (size_t)bh * S * row_bytes;
                const char* vb = (const char*)v_in + (size_t)bh * S * row_bytes;
                for (int j = 0; j < sink; ++j) rows[j] = j;
                if (k_sel > 0) {
                    row_norms(kb, dtype, D, lo, hi, score);
                    if (mode == MODE_SNAPKV) snapkv_scores(score, score + R, R, dtype, pool_kernel);
                    const int descending = (mode == MODE_HIGH || mode == MODE_SNAPKV);
                    for (int i = 0; i < R; ++i) { order[i].key = descending ? -score[i] : score[i]; order[i].idx = i; }
                    qsort(order, (size_t)R, sizeof(keyed), cmp_keyed);          /* full argsort */
                    for (int i = 0; i < k_sel; ++i) rows[sink + i] = lo + order[i].idx; /* [:k] */
                    qsort(rows + sink, (size_t)k_sel, sizeof(int32_t), cmp_i32);  /* temporal order */
                }
                for (int j = 0; j < tail; ++j) rows[sink + k_sel + j] = S - tail + j;
                char* ko = (char*)k_out + (size_t)bh * C * row_bytes;
                char* vo = (char*)v_out + (size_t)bh * C * row_bytes;
                for (int j = 0; j < C; ++j) {
                    memcpy(ko + (size_t)j * row_bytes, kb + (size_t)rows[j] * row_bytes, row_bytes);
                    memcpy(vo + (size_t)j * row_bytes, vb + (size_t)rows[j] * row_bytes, row_bytes);
                }
                if (rows_out) memcpy(rows_out + (size_t)bh * C, rows, sizeof(int32_t) * (size_t)C);
            }
        }
        free(score);
        free(order);
        free(rows);
    }
    return failed ? -1 : 0;
}

/* norms only: out[b,h,r] as float, r in [lo,hi) */
int kvc_oracle_norms(int dtype, int B, int H, int S, int D, const void* k_in, int lo, int hi, float* out,
                     int nthreads) {
    if (dtype < 0 || dtype > 2 || lo < 0 || hi > S || hi < lo) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const size_t row_bytes = (size_t)D * elem_size(dtype);
#pragma omp parallel for schedule(static)
    for (int bh = 0; bh < B * H; ++bh)
        row_norms((const char*)k_in + (size_t)bh * S * row_bytes, dtype, D, lo, hi, out + (size_t)bh * (hi - lo));
    return 0;
}

int kvc_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
