/* ORACLE (test infrastructure, never shipped as the product): scalar C restatement of the frozen
 * lifting spec, SURVEY.md Appendix A (steps a-1..a-3), and of the sequential summation order of
 * torch_scatter.scatter_mean on CPU (step a-4).
 *
 * PARITY UNPINNED for a-1..a-3: the reference has no code for projection / visibility / bilinear
 * gather / view mean (consumer only: /root/reference/segdino3d/datasets/dataset/scannet200.py:219-234).
 * This file exists (1) as an independent cross-check of oracle/lift_oracle.py (vectorised torch) -- the
 * two must agree bit for bit, tests/test_oracle.py -- and (2) as the multi-threaded CPU baseline that
 * bench.py times beside the GPU path (cpu_baseline.kind = "port").
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp): contraction OFF so that every
 * mul/add below rounds separately, exactly as Appendix A demands.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static inline float bf16_to_float(uint16_t h) {
    uint32_t bits = (uint32_t)h << 16;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* fmap_dtype: 0 = f32, 1 = f16, 2 = bf16.   depth_dtype: 0 = f32 metres, 1 = u16 millimetres. */
static inline float load_feat(const void* fmap, int dtype, int64_t i) {
    if (dtype == 0) return ((const float*)fmap)[i];
    if (dtype == 1) return half_to_float(((const uint16_t*)fmap)[i]);
    return bf16_to_float(((const uint16_t*)fmap)[i]);
}

int sd3d_ref_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* OpenMP team size of the following calls, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1) */
void sd3d_ref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Appendix A, verbatim loop nest. Outputs: sum[N,C], count[N]; pix_idx[V,N] / vis[V,N] optional (NULL). */
void sd3d_ref_lift(const float* xyz, int64_t N, const float* K4, const float* w2c, int V, const void* depth,
                   int depth_dtype, int Hd, int Wd, const void* fmap, int fmap_dtype, int Hl, int Wl, int C,
                   float s, float tau, float z_near, float* sum, int32_t* count, int32_t* pix_idx,
                   uint8_t* vis) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t p = 0; p < N; ++p) {
        const float x = xyz[3 * p], y = xyz[3 * p + 1], z = xyz[3 * p + 2];
        float* acc = sum + p * (int64_t)C;
        for (int c = 0; c < C; ++c) acc[c] = 0.0f;
        int32_t cnt = 0;
        for (int v = 0; v < V; ++v) {
            if (pix_idx) pix_idx[(int64_t)v * N + p] = -1;
            if (vis) vis[(int64_t)v * N + p] = 0;
            const float* R = w2c + 12 * v;
            const float xc = ((R[0] * x + R[1] * y) + R[2] * z) + R[3];
            const float yc = ((R[4] * x + R[5] * y) + R[6] * z) + R[7];
            const float zc = ((R[8] * x + R[9] * y) + R[10] * z) + R[11];
            if (!(zc > z_near)) continue;
            const float fx = K4[4 * v], fy = K4[4 * v + 1], cx = K4[4 * v + 2], cy = K4[4 * v + 3];
            const float u = (fx * xc) / zc + cx;
            const float w = (fy * yc) / zc + cy;
            const float uif = floorf(u + 0.5f), wif = floorf(w + 0.5f);
            if (!(uif >= 0.0f && uif < (float)Wd && wif >= 0.0f && wif < (float)Hd)) continue;
            const int ui = (int)uif, wi = (int)wif;
            const int64_t pix = (int64_t)wi * Wd + ui;
            float d;
            if (depth_dtype == 0) d = ((const float*)depth)[(int64_t)v * Hd * Wd + pix];
            else d = (float)((const uint16_t*)depth)[(int64_t)v * Hd * Wd + pix] * 0.001f;
            if (!(d > 0.0f)) continue;
            if (!(fabsf(d - zc) <= tau)) continue;
            if (pix_idx) pix_idx[(int64_t)v * N + p] = (int32_t)pix;
            if (vis) vis[(int64_t)v * N + p] = 1;
            cnt += 1;
            const float uf = (u + 0.5f) / s - 0.5f, wf = (w + 0.5f) / s - 0.5f;
            const float x0f = floorf(uf), y0f = floorf(wf);
            const int x0 = (int)x0f, y0 = (int)y0f;
            const float ax = uf - x0f, ay = wf - y0f;
            const float w00 = (1.0f - ax) * (1.0f - ay), w01 = ax * (1.0f - ay);
            const float w10 = (1.0f - ax) * ay, w11 = ax * ay;
            const int okx0 = x0 >= 0 && x0 < Wl, okx1 = x0 + 1 >= 0 && x0 + 1 < Wl;
            const int oky0 = y0 >= 0 && y0 < Hl, oky1 = y0 + 1 >= 0 && y0 + 1 < Hl;
            const int64_t base = (int64_t)v * Hl * Wl;
            const int64_t i00 = (base + (int64_t)y0 * Wl + x0) * C, i01 = i00 + C;
            const int64_t i10 = i00 + (int64_t)Wl * C, i11 = i10 + C;
            for (int c = 0; c < C; ++c) {
                const float t00 = (oky0 && okx0) ? load_feat(fmap, fmap_dtype, i00 + c) : 0.0f;
                const float t01 = (oky0 && okx1) ? load_feat(fmap, fmap_dtype, i01 + c) : 0.0f;
                const float t10 = (oky1 && okx0) ? load_feat(fmap, fmap_dtype, i10 + c) : 0.0f;
                const float t11 = (oky1 && okx1) ? load_feat(fmap, fmap_dtype, i11 + c) : 0.0f;
                const float f = ((w00 * t00 + w01 * t01) + w10 * t10) + w11 * t11;
                acc[c] = acc[c] + f;
            }
        }
        count[p] = cnt;
    }
}

/* feat = sum / (float)max(count,1), in place (Appendix A `feat_l[p,c]`). */
void sd3d_ref_finalize(float* sum, const int32_t* count, int64_t N, int C) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < N; ++p) {
        const float d = (float)(count[p] > 1 ? count[p] : 1);
        for (int c = 0; c < C; ++c) sum[p * (int64_t)C + c] = sum[p * (int64_t)C + c] / d;
    }
}

/* scatter_mean(src[N,C], idx[N], dim=0) with the CPU summation order of aten scatter_add_
 * (ascending point index per destination row). out[S,C] is fully overwritten. Single pass,
 * sequential over points (that IS the order); rows of different superpoints are independent. */
void sd3d_ref_scatter_mean(const float* src, const int64_t* idx, int64_t N, int64_t S, int C, float* out) {
    for (int64_t i = 0; i < S * (int64_t)C; ++i) out[i] = 0.0f;
    /* column-block parallel: every thread walks all points in ascending order for its channels */
#pragma omp parallel
    {
#ifdef _OPENMP
        const int nt = omp_get_num_threads(), t = omp_get_thread_num();
#else
        const int nt = 1, t = 0;
#endif
        const int c0 = (int)((int64_t)C * t / nt), c1 = (int)((int64_t)C * (t + 1) / nt);
        if (c1 > c0) {
            for (int64_t p = 0; p < N; ++p) {
                const int64_t s = idx[p];
                if (s < 0 || s >= S) continue;
                float* o = out + s * (int64_t)C;
                const float* r = src + p * (int64_t)C;
                for (int c = c0; c < c1; ++c) o[c] = o[c] + r[c];
            }
        }
    }
    /* counts are accumulated in src dtype (float) by torch_scatter: ones.scatter_add_ ; clamp >= 1 */
    for (int64_t s0 = 0; s0 < S; s0 += 4096) {
        float c4096[4096];
        const int64_t s1 = s0 + 4096 < S ? s0 + 4096 : S;
        for (int64_t s = s0; s < s1; ++s) c4096[s - s0] = 0.0f;
        for (int64_t p = 0; p < N; ++p) {
            const int64_t s = idx[p];
            if (s >= s0 && s < s1) c4096[s - s0] = c4096[s - s0] + 1.0f;
        }
        for (int64_t s = s0; s < s1; ++s) {
            const float d = c4096[s - s0] < 1.0f ? 1.0f : c4096[s - s0];
            for (int c = 0; c < C; ++c) out[s * (int64_t)C + c] = out[s * (int64_t)C + c] / d;
        }
    }
}
