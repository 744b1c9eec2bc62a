/* kjarni_oracle.c -- C restatement of the reference's CPU algorithm for the hot path
 * (encoder forward + pooling, cosine scan), structured like the reference's own CPU kernels
 * so that it is a fair timed baseline on the GPU box's host cores.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (checked against the pinned numpy oracle,
 * oracle/kjarni_oracle.py, and through it against the reference's golden vectors) and by
 * bench.py's cpu_baseline / --impl reference legs.  Nothing under kjarni_b200/ links or loads it.
 *
 * Parity status: PINNED via tests/test_oracle_c.py (max-abs <= 2e-5 against the numpy oracle, which is
 * itself pinned to the reference's goldens).  The reference (Rust) cannot be compiled in this image.
 *
 * What it follows (file:line into /root/reference/crates/, KT = kjarni-transformers/src, KR = kjarni-rag/src):
 *   GEMM        KT/cpu/ops/matmul.rs:370-479 (rayon over 64-token blocks) -> KT/cpu/kernels/x86/f32.rs:9-124
 *               (4 tokens x 3 outputs, 8-lane AVX2 FMA partial sums, horizontal sum, bias added once)
 *   QKV         KT/cpu/encoder/qkv_projection.rs:93-138
 *   attention   KT/cpu/encoder/encoder_self_attention.rs:143-307 (scale :244-249, mask :311-325, merge :384-425)
 *   softmax     KT/activations.rs:223-242,259-279 (max-subtract, divide only if sum > 0)
 *   GELU        KT/activations.rs:57-59 (0.5 x (1 + erff(x / sqrt2)))
 *   LayerNorm   KT/cpu/normalization/layer_norm.rs:37-134 (biased variance, eps inside the sqrt)
 *   layer       KT/cpu/encoder/encoder_layer.rs:113-179 (post-norm)
 *   embeddings  KT/cpu/embeddings/mod.rs:181-326
 *   pooling     KT/pooling/mod.rs:11-34, KT/cpu/encoder/traits.rs:529-536
 *   scan        KR/segment.rs:307-337,355-370 (per-row cosine with cached norms, select top-k, sort)
 * Threads: OpenMP over the same block decomposition the reference hands to rayon.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    const float *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *ln1_g, *ln1_b, *w1, *b1, *w2, *b2, *ln2_g, *ln2_b;
} KoLayer;

typedef struct {
    int hidden, layers, heads, inter, vocab, max_pos, type_vocab, pos_offset;
    float eps;
    const float *word, *pos, *type, *emb_g, *emb_b;
    const KoLayer* layer;
} KoModel;

/* rayon::ThreadPoolBuilder::num_threads(physical_cores) in kjarni_init (kjarni-ffi/src/lib.rs:36-40) */
void ko_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ko_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline float hsum8(__m256 v) {
    __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
    lo = _mm_add_ps(lo, hi);
    lo = _mm_hadd_ps(lo, lo);
    lo = _mm_hadd_ps(lo, lo);
    return _mm_cvtss_f32(lo);
}

static inline float dot_f32(const float* a, const float* b, int k) {
    __m256 acc = _mm256_setzero_ps();
    int i = 0;
    for (; i + 8 <= k; i += 8) acc = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), acc);
    float s = hsum8(acc);
    for (; i < k; ++i) s += a[i] * b[i];
    return s;
}

/* 4 tokens x 3 outputs micro-kernel (matmul_block_4x3_f32). */
static void block_4x3(const float* a, int lda, const float* w, int ldw, int k, float* c, int ldc, const float* bias, int col) {
    __m256 acc[4][3];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 3; ++j) acc[i][j] = _mm256_setzero_ps();
    int kk = 0;
    for (; kk + 8 <= k; kk += 8) {
        const __m256 w0 = _mm256_loadu_ps(w + kk), w1 = _mm256_loadu_ps(w + ldw + kk), w2 = _mm256_loadu_ps(w + 2 * ldw + kk);
        for (int i = 0; i < 4; ++i) {
            const __m256 av = _mm256_loadu_ps(a + i * lda + kk);
            acc[i][0] = _mm256_fmadd_ps(av, w0, acc[i][0]);
            acc[i][1] = _mm256_fmadd_ps(av, w1, acc[i][1]);
            acc[i][2] = _mm256_fmadd_ps(av, w2, acc[i][2]);
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = hsum8(acc[i][j]);
            for (int t = kk; t < k; ++t) s += a[i * lda + t] * w[j * ldw + t];
            c[i * ldc + j] = s + (bias ? bias[col + j] : 0.0f);
        }
}

/* C[M,N] = A[M,K] W[N,K]^T + b ; parallel over 64-token blocks. */
void ko_linear(const float* a, const float* w, const float* bias, float* c, int m, int n, int k) {
    const int blocks = (m + 63) / 64;
#pragma omp parallel for schedule(dynamic, 1)
    for (int blk = 0; blk < blocks; ++blk) {
        const int r0 = blk * 64, r1 = r0 + 64 < m ? r0 + 64 : m;
        int r = r0;
        for (; r + 4 <= r1; r += 4) {
            int j = 0;
            for (; j + 3 <= n; j += 3) block_4x3(a + (size_t)r * k, k, w + (size_t)j * k, k, k, c + (size_t)r * n + j, n, bias, j);
            for (; j < n; ++j)
                for (int i = 0; i < 4; ++i) c[(size_t)(r + i) * n + j] = dot_f32(a + (size_t)(r + i) * k, w + (size_t)j * k, k) + (bias ? bias[j] : 0.0f);
        }
        for (; r < r1; ++r)
            for (int j = 0; j < n; ++j) c[(size_t)r * n + j] = dot_f32(a + (size_t)r * k, w + (size_t)j * k, k) + (bias ? bias[j] : 0.0f);
    }
}

void ko_layer_norm(float* x, const float* g, const float* b, float eps, int m, int h) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < m; ++r) {
        float* p = x + (size_t)r * h;
        float mean = 0.f;
        for (int i = 0; i < h; ++i) mean += p[i];
        mean /= (float)h;
        float var = 0.f;
        for (int i = 0; i < h; ++i) { const float d = p[i] - mean; var += d * d; }
        var /= (float)h;
        const float rstd = 1.0f / sqrtf(var + eps);
        for (int i = 0; i < h; ++i) p[i] = (p[i] - mean) * rstd * g[i] + b[i];
    }
}

static void softmax_row(float* s, int n) {
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i) mx = s[i] > mx ? s[i] : mx;
    float sum = 0.f;
    for (int i = 0; i < n; ++i) { s[i] = expf(s[i] - mx); sum += s[i]; }
    if (sum > 0.f) {
        const float inv = 1.0f / sum;
        for (int i = 0; i < n; ++i) s[i] *= inv;
    }
}

/* q,k,v [B*S,H] -> ctx [B*S,H]; parallel over (batch, head). mask value: -inf (noalloc) or -1e9. */
static void attention(const float* q, const float* k, const float* v, const float* mask, float* ctx, int B, int S, int H, int heads, int noalloc) {
    const int d = H / heads;
    const float scale = 1.0f / sqrtf((float)d);
    const float fill = noalloc ? -INFINITY : -1e9f;
#pragma omp parallel
    {
        float* sc = (float*)malloc((size_t)S * sizeof(float));
#pragma omp for schedule(dynamic, 1) collapse(2)
        for (int b = 0; b < B; ++b)
            for (int h = 0; h < heads; ++h) {
                for (int i = 0; i < S; ++i) {
                    const float* qi = q + ((size_t)(b * S + i)) * H + h * d;
                    for (int j = 0; j < S; ++j) {
                        const float* kj = k + ((size_t)(b * S + j)) * H + h * d;
                        float s = dot_f32(qi, kj, d) * scale;
                        if (mask && mask[b * S + j] == 0.0f) s = fill;
                        sc[j] = s;
                    }
                    softmax_row(sc, S);
                    float* o = ctx + ((size_t)(b * S + i)) * H + h * d;
                    for (int t = 0; t < d; ++t) o[t] = 0.f;
                    for (int j = 0; j < S; ++j) {
                        const float pj = sc[j];
                        const float* vj = v + ((size_t)(b * S + j)) * H + h * d;
                        for (int t = 0; t < d; ++t) o[t] += pj * vj[t];
                    }
                }
            }
        free(sc);
    }
}

/* ids u32 [B,S], mask f32 [B,S] (or NULL), type_ids u32 [B,S] (or NULL) -> hidden f32 [B,S,H] */
int ko_encoder_forward(const KoModel* m, const uint32_t* ids, const float* mask, const uint32_t* type_ids, int B, int S, int noalloc, float* hidden) {
    const int H = m->hidden, I = m->inter, M = B * S;
    float* x = hidden;
    float* q = (float*)malloc((size_t)M * H * 4);
    float* k = (float*)malloc((size_t)M * H * 4);
    float* v = (float*)malloc((size_t)M * H * 4);
    float* ctx = (float*)malloc((size_t)M * H * 4);
    float* t = (float*)malloc((size_t)M * (I > H ? I : H) * 4);
    float* y = (float*)malloc((size_t)M * H * 4);
    if (!q || !k || !v || !ctx || !t || !y) return -1;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < M; ++r) {
        const int s = r % S;
        float* p = x + (size_t)r * H;
        const uint32_t id = ids[r];
        if ((int)id < m->vocab) memcpy(p, m->word + (size_t)id * H, (size_t)H * 4);
        else memset(p, 0, (size_t)H * 4);
        const int pi = s + m->pos_offset;
        if (pi < m->max_pos) { const float* pp = m->pos + (size_t)pi * H; for (int i = 0; i < H; ++i) p[i] += pp[i]; }
        if (m->type) { const float* tp = m->type + (size_t)(type_ids ? type_ids[r] : 0) * H; for (int i = 0; i < H; ++i) p[i] += tp[i]; }
    }
    ko_layer_norm(x, m->emb_g, m->emb_b, m->eps, M, H);
    for (int l = 0; l < m->layers; ++l) {
        const KoLayer* L = &m->layer[l];
        ko_linear(x, L->wq, L->bq, q, M, H, H);
        ko_linear(x, L->wk, L->bk, k, M, H, H);
        ko_linear(x, L->wv, L->bv, v, M, H, H);
        attention(q, k, v, mask, ctx, B, S, H, m->heads, noalloc);
        ko_linear(ctx, L->wo, L->bo, y, M, H, H);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < (size_t)M * H; ++i) x[i] += y[i];
        ko_layer_norm(x, L->ln1_g, L->ln1_b, m->eps, M, H);
        ko_linear(x, L->w1, L->b1, t, M, I, H);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < (size_t)M * I; ++i) t[i] = 0.5f * t[i] * (1.0f + erff(t[i] * 0.70710678118654752f));
        ko_linear(t, L->w2, L->b2, y, M, H, I);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < (size_t)M * H; ++i) x[i] += y[i];
        ko_layer_norm(x, L->ln2_g, L->ln2_b, m->eps, M, H);
    }
    free(q); free(k); free(v); free(ctx); free(t); free(y);
    return 0;
}

/* masked mean pool (count 0 -> token-0 row) + optional L2 normalise: hidden [B,S,H] -> out [B,H] */
void ko_mean_pool_l2(const float* hidden, const float* mask, int B, int S, int H, int normalize, float* out) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        float* o = out + (size_t)b * H;
        for (int i = 0; i < H; ++i) o[i] = 0.f;
        float cnt = 0.f;
        for (int s = 0; s < S; ++s) {
            const float mk = mask ? mask[b * S + s] : 1.0f;
            cnt += mk;
            const float* p = hidden + ((size_t)(b * S + s)) * H;
            for (int i = 0; i < H; ++i) o[i] += p[i] * mk;
        }
        if (cnt > 0.f) for (int i = 0; i < H; ++i) o[i] /= cnt;
        else memcpy(o, hidden + (size_t)b * S * H, (size_t)H * 4);
        if (normalize) {
            float n2 = 0.f;
            for (int i = 0; i < H; ++i) n2 += o[i] * o[i];
            const float nm = sqrtf(n2);
            if (nm > 0.f) for (int i = 0; i < H; ++i) o[i] /= nm;
        }
    }
}

int ko_embed(const KoModel* m, const uint32_t* ids, const float* mask, int B, int S, int noalloc, float* out) {
    float* hidden = (float*)malloc((size_t)B * S * m->hidden * 4);
    if (!hidden) return -1;
    const int rc = ko_encoder_forward(m, ids, mask, NULL, B, S, noalloc, hidden);
    if (rc == 0) ko_mean_pool_l2(hidden, mask, B, S, m->hidden, 1, out);
    free(hidden);
    return rc;
}

/* Segment::search_vectors over rows [n,dim] with cached norms, for nq queries: per query a scalar pass over every row
 * (sequential fp32 accumulation like the reference), top-k by (score desc, id asc).  Parallel over queries. */
void ko_scan_topk(const float* rows, const float* norms, uint64_t n, int dim, const float* queries, int nq, int k, uint64_t* out_ids,
                  float* out_scores) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int qi = 0; qi < nq; ++qi) {
        const float* q = queries + (size_t)qi * dim;
        float qn = 0.f;
        for (int i = 0; i < dim; ++i) qn += q[i] * q[i];
        qn = sqrtf(qn);
        uint64_t* oi = out_ids + (size_t)qi * k;
        float* os = out_scores + (size_t)qi * k;
        int cnt = 0;
        for (int j = 0; j < k; ++j) { oi[j] = UINT64_MAX; os[j] = -INFINITY; }
        if (qn < 1e-9f) continue;
        for (uint64_t r = 0; r < n; ++r) {
            const float rn = norms[r];
            const float s = rn < 1e-9f ? 0.0f : dot_f32(q, rows + r * dim, dim) / (qn * rn);
            if (cnt < k || s > os[k - 1]) {  /* insert keeping (score desc, id asc): equal scores stay behind earlier ids */
                int pos = cnt < k ? cnt : k - 1;
                while (pos > 0 && os[pos - 1] < s) { os[pos] = os[pos - 1]; oi[pos] = oi[pos - 1]; --pos; }
                os[pos] = s;
                oi[pos] = r;
                if (cnt < k) ++cnt;
            }
        }
    }
}

void ko_row_norms(const float* rows, uint64_t n, int dim, float* norms) {
#pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < n; ++r) {
        float s = 0.f;
        const float* p = rows + r * dim;
        for (int i = 0; i < dim; ++i) s += p[i] * p[i];
        norms[r] = sqrtf(s);
    }
}
