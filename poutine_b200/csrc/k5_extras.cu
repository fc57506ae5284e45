// K5 — two per-site statistics the north_star names next to the live path, outside the reference's LIVE code:
//
//   * exact point-wise null: the m -> infinity limit of the reference's Monte-Carlo estimate
//     r/m (calc_pointwise_estimate_binom, HC.java:4812-4838).  Under a uniform permutation of the pheno labels
//     (permute_phenos, HC.java:4890-4923) the allele's n sampled leaves draw j missing labels ~ Hyper(P, n_miss, n) and,
//     given j, k cases ~ Hyper(P - n_miss, n_case, n - j); the replicate counts iff f(n - j, k) <= p_obs (HC.java:4239).
//   * Fisher's exact test, two-sided, of the 2x2 table {a1case, a1ctrl; a2case, a2ctrl}: the reference's DEAD
//     fishers_exact path (HC.java:5008-5093) pipes these tables to an R script that is not in the repository; the
//     definition here is R's fisher.test: sum of the table probabilities <= P(observed) * (1 + 1e-7).
//
// PARITY UNPINNED: neither has a reference implementation that can run; both are checked against
// exact rational arithmetic in tests/ (relative tolerance 1e-12).  Hypergeometric weights come from the ratio
// recurrence started at the mode and are normalised by their own sum — no lgamma, a few ulp per step.
#include <algorithm>

#include "pb_internal.h"

namespace {

// X ~ Hyper(pop, succ, draws): w(x+1)/w(x) = (succ-x)(draws-x) / ((x+1)(pop-succ-draws+x+1)), support [lo, hi]
struct Hyper {
    int64_t pop, succ, draws, lo, hi, mode;
    __device__ Hyper(int64_t pop_, int64_t succ_, int64_t draws_) : pop(pop_), succ(succ_), draws(draws_) {
        lo = max((int64_t)0, draws - (pop - succ));
        hi = min(draws, succ);
        mode = ((draws + 1) * (succ + 1)) / (pop + 2);
        mode = min(max(mode, lo), hi);
    }
    __device__ double up(int64_t x) const {   // w(x+1) / w(x)
        return ((double)(succ - x) * (double)(draws - x)) / ((double)(x + 1) * (double)(pop - succ - draws + x + 1));
    }
};

// sum of the normalised weights over {x : accept(x)}
template <typename F>
__device__ double hyper_mass(const Hyper& h, F accept) {
    double z = 0.0, acc = 0.0, w = 1.0;
    for (int64_t x = h.mode; x <= h.hi; x++) {
        z += w;
        if (accept(x)) acc += w;
        if (x < h.hi) w *= h.up(x);
    }
    w = 1.0;
    for (int64_t x = h.mode - 1; x >= h.lo; x--) {
        w /= h.up(x);
        z += w;
        if (accept(x)) acc += w;
    }
    return acc / z;
}

__global__ void k5_exact_pointwise_kernel(const AlleleMeta* __restrict__ alleles, int64_t n_slots, const double* __restrict__ pobs,
                                          const double* __restrict__ tab, int32_t P, int32_t n_case, int32_t n_miss,
                                          double* __restrict__ out) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_slots) return;
    const AlleleMeta m = alleles[a];
    if (!(m.flags & PB_AF_IN_USE)) { out[a] = __longlong_as_double(0x7FF8000000000000ll); return; }
    const double po = pobs[a];
    const int64_t n = m.n;
    const Hyper hj(P, n_miss, n);            // j = how many of the allele's n sampled leaves draw a missing label
    double zj = 0.0, accj = 0.0, wj = 1.0;
    auto inner = [&](int64_t j) {
        const int64_t nr = n - j;
        const double* row = tab + nr * (nr + 1) / 2;
        const Hyper hk(P - n_miss, n_case, nr);
        return hyper_mass(hk, [&](int64_t k) { return row[k] <= po; });
    };
    for (int64_t j = hj.mode; j <= hj.hi; j++) {
        zj += wj; accj += wj * inner(j);
        if (j < hj.hi) wj *= hj.up(j);
    }
    wj = 1.0;
    for (int64_t j = hj.mode - 1; j >= hj.lo; j--) {
        wj /= hj.up(j);
        zj += wj; accj += wj * inner(j);
    }
    out[a] = accj / zj;
}

__global__ void k5_fisher_kernel(const AlleleMeta* __restrict__ alleles, int64_t S_inf, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S_inf) return;
    const AlleleMeta m1 = alleles[2 * i], m2 = alleles[2 * i + 1];
    const int64_t a = m1.obs_case, b = m1.obs_ctrl, c = m2.obs_case, d = m2.obs_ctrl;
    // rows m = a+b (allele 1), n = c+d (allele 2); first column k = a+c (cases); x = a ~ Hyper(m+n, m, k)
    const Hyper h(a + b + c + d, a + b, a + c);
    // weight of the observed table relative to the mode
    double wobs = 1.0;
    if (a >= h.mode) for (int64_t x = h.mode; x < a; x++) wobs *= h.up(x);
    else for (int64_t x = h.mode - 1; x >= a; x--) wobs /= h.up(x);
    const double thr = wobs * (1.0 + 1e-7);   // R: relErr <- 1 + 10^-7
    double z = 0.0, acc = 0.0, w = 1.0;
    for (int64_t x = h.mode; x <= h.hi; x++) {
        z += w;
        if (w <= thr) acc += w;
        if (x < h.hi) w *= h.up(x);
    }
    w = 1.0;
    for (int64_t x = h.mode - 1; x >= h.lo; x--) {
        w /= h.up(x);
        z += w;
        if (w <= thr) acc += w;
    }
    out[i] = acc / z;
}

}  // namespace

int pbk_extras(pb_ctx* ctx, double* h_exact_pointwise, double* h_fisher) {
    const int64_t S_inf = ctx->S_inf;
    if (S_inf == 0) return PB_OK;
    if (3 * S_inf > ctx->extras_cap) {
        if (ctx->d_extras) { cudaFree(ctx->d_extras); ctx->d_extras = nullptr; ctx->extras_cap = 0; }
        PB_CUDA(cudaMalloc(&ctx->d_extras, (size_t)(3 * S_inf) * 8));
        ctx->extras_cap = 3 * S_inf;
    }
    const int T = 128;
    cudaStream_t st = ctx->stream;
    if (h_exact_pointwise) {
        k5_exact_pointwise_kernel<<<(unsigned)((2 * S_inf + T - 1) / T), T, 0, st>>>(ctx->d_alleles, 2 * S_inf, ctx->d_pobs, ctx->d_ptable, ctx->P,
                                                                                  ctx->n_case, ctx->n_miss, ctx->d_extras);
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaMemcpyAsync(h_exact_pointwise, ctx->d_extras, (size_t)(2 * S_inf) * 8, cudaMemcpyDeviceToHost, st));
    }
    if (h_fisher) {
        k5_fisher_kernel<<<(unsigned)((S_inf + T - 1) / T), T, 0, st>>>(ctx->d_alleles, S_inf, ctx->d_extras + 2 * S_inf);
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaMemcpyAsync(h_fisher, ctx->d_extras + 2 * S_inf, (size_t)S_inf * 8, cudaMemcpyDeviceToHost, st));
    }
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}
