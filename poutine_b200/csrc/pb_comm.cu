// The hot path's one exchange, over NVLink peer memory (sm_100a, NVSwitch box).
//
// Site-sharded contexts need the per-replicate minimum p-value over ALL sites (the reference takes it over the whole
// alignment: HC.java:3239-3243, pooled for a1 and a2).  Instead of a library collective between K3 and K4, every rank
// exposes its maxT[m] buffer to its peers and the sort kernel of K4 (k4_pvalues.cu) takes the minimum over all ranks with
// peer loads while it fills its tiles: compute step and collective are ONE kernel.
//
// Protocol (see pb_xchg in pb_internal.h): block = {ready flag of rank q at word q, error word at 64, maxT buffer 0,
// maxT buffer 1}.  Epoch e (one per pb_run_assoc; every rank runs the same number of them):
//   K3 of rank r min-es into buffer e & 1 of its own block
//   xchg_signal_kernel (same stream, after K3): st.release.sys  block_q.flag[r] = e   for every peer q
//   k4_sort_coop_kernel: every block waits until flag[q] >= e for all q (ld.acquire.sys on its OWN block), then
//                        loads min_q block_q.buffer[e & 1][i] (ld.relaxed.sys over NVLink) instead of its own maxT[i]
// A peer can be at most one epoch ahead (it needs this rank's flag of e + 1 to pass its own wait), and epoch e + 1 writes
// the other buffer, so no "done reading" handshake is needed.  The wait has a 20 s time-out (a peer that died): the error
// word is set, the kernel carries on, and pb_run_assoc returns PB_ERR_NCCL.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "pb_internal.h"

__global__ void xchg_signal_kernel(unsigned long long* const* __restrict__ peers, int world, int rank, unsigned long long epoch) {
    const int q = threadIdx.x;
    if (q < world) {
        __threadfence_system();
        unsigned long long* f = peers[q] + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    }
}

static int alloc_block(pb_ctx* ctx) {
    pb_xchg& x = ctx->xchg;
    if (x.block) return PB_OK;
    const int64_t m = ctx->cfg.n_replicates;
    x.block_bytes = (size_t)(PB_XCHG_HDR_WORDS + 2 * m) * 8;
    cudaError_t e = cudaMalloc(&x.block, x.block_bytes);
    if (e != cudaSuccess) { x.block = nullptr; ctx->err = std::string("cudaMalloc of the exchange block: ") + cudaGetErrorString(e); return PB_ERR_OOM; }
    PB_CUDA(cudaMemset(x.block, 0, x.block_bytes));
    PB_CUDA(cudaMalloc(&x.d_pooled, (size_t)m * 8));
    PB_CUDA(cudaMalloc(&x.d_peer_tbl, PB_XCHG_MAX_WORLD * sizeof(void*)));
    PB_CUDA(cudaDeviceSynchronize());      // the zeroed flags must be in place before any peer can map and write them
    return PB_OK;
}

static int finish_connect(pb_ctx* ctx, int rank, int world) {
    pb_xchg& x = ctx->xchg;
    PB_CUDA(cudaMemcpy(x.d_peer_tbl, x.peer, PB_XCHG_MAX_WORLD * sizeof(void*), cudaMemcpyHostToDevice));
    ctx->rank = rank; ctx->world = world;
    x.epoch = 0;
    x.active = world > 1;
    return PB_OK;
}

int pbx_export(pb_ctx* ctx, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "pb_comm_export hands out 64 bytes");
    int rc = alloc_block(ctx);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    PB_CUDA(cudaIpcGetMemHandle(&h, ctx->xchg.block));
    memcpy(handle64, &h, 64);
    return PB_OK;
}

int pbx_connect_ipc(pb_ctx* ctx, const void* handles, int rank, int world) {
    if (world > PB_XCHG_MAX_WORLD) { ctx->err = "pb_comm_connect: at most 64 ranks"; return PB_ERR_INVALID; }
    int rc = alloc_block(ctx);
    if (rc) return rc;
    pb_xchg& x = ctx->xchg;
    for (int q = 0; q < world; q++) {
        if (q == rank) { x.peer[q] = x.block; x.peer_ipc[q] = false; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * (size_t)q, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            ctx->err = std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(q) + "): " + cudaGetErrorString(e) +
                       " — peer memory is not reachable from this process; use pb_comm_init (NCCL) instead";
            cudaGetLastError();
            return PB_ERR_NCCL;
        }
        x.peer[q] = (unsigned long long*)p;
        x.peer_ipc[q] = true;
    }
    return finish_connect(ctx, rank, world);
}

int pbx_connect_local(pb_ctx** ctxs, int n) {
    if (n > PB_XCHG_MAX_WORLD) { ctxs[0]->err = "pb_group: at most 64 devices"; return PB_ERR_INVALID; }
    for (int i = 0; i < n; i++) {
        pb_ctx* ctx = ctxs[i];
        PB_CUDA(cudaSetDevice(ctx->device));
        int rc = alloc_block(ctx);
        if (rc) return rc;
        for (int j = 0; j < n; j++) {
            if (j == i || ctxs[j]->device == ctx->device) continue;
            int can = 0;
            PB_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[j]->device));
            if (!can) {
                ctx->err = "pb_group: device " + std::to_string(ctx->device) + " cannot access device " + std::to_string(ctxs[j]->device) +
                           " (no NVLink / PCIe peer path): the shards' exchange needs peer memory";
                return PB_ERR_NCCL;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                return PB_ERR_CUDA;
            }
            cudaGetLastError();
        }
    }
    for (int i = 0; i < n; i++) {
        pb_ctx* ctx = ctxs[i];
        PB_CUDA(cudaSetDevice(ctx->device));
        for (int j = 0; j < n; j++) { ctx->xchg.peer[j] = ctxs[j]->xchg.block; ctx->xchg.peer_ipc[j] = false; }
        int rc = finish_connect(ctx, i, n);
        if (rc) return rc;
    }
    return PB_OK;
}

int pbx_begin_epoch(pb_ctx* ctx) {
    pb_xchg& x = ctx->xchg;
    if (!x.active) return PB_OK;
    x.epoch++;
    ctx->d_maxT = x.block + PB_XCHG_HDR_WORDS + (int64_t)(x.epoch & 1ull) * ctx->cfg.n_replicates;
    return PB_OK;
}

int pbx_signal(pb_ctx* ctx) {
    pb_xchg& x = ctx->xchg;
    if (!x.active) return PB_OK;
    xchg_signal_kernel<<<1, PB_XCHG_MAX_WORLD, 0, ctx->stream>>>(x.d_peer_tbl, ctx->world, ctx->rank, x.epoch);
    ctx->tm.n_launches += 1;
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

XchgView pbx_view(pb_ctx* ctx) {
    XchgView v{};
    const pb_xchg& x = ctx->xchg;
    if (!x.active) { v.peers = nullptr; v.pooled = nullptr; v.world = 1; v.rank = 0; v.epoch = 0; v.buf_off = 0; return v; }
    v.peers = x.d_peer_tbl; v.pooled = x.d_pooled; v.world = ctx->world; v.rank = ctx->rank; v.epoch = x.epoch;
    v.buf_off = PB_XCHG_HDR_WORDS + (long long)(x.epoch & 1ull) * ctx->cfg.n_replicates;
    return v;
}

// ---- label sharing ---------------------------------------------------------------------------------------------
// Every site shard sweeps the SAME label bit-matrix (same seed, same permutations).  Instead of every rank drawing all
// P x m labels, rank r draws the replicate columns of generator blocks [T*r/G, T*(r+1)/G) of a tile into its own matrix,
// stores the tile's epoch into every peer's "slice ready" flag, and pulls the other ranks' column slices out of their
// matrices with 16-byte NVLink loads (k3_gather_kernel) — the generator's integer work drops by the number of ranks and the
// matrix crosses NVLink once.  A buffer is redrawn only after every peer has acknowledged pulling its previous content
// ("slice pulled" flags).  Same 20 s time-out / error word as the maxT exchange.
__device__ __forceinline__ unsigned long long ldx_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_flag(const unsigned long long* f, unsigned long long epoch, unsigned long long* err) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ldx_acquire_sys(f) < epoch) {
        __nanosleep(100);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) { atomicExch(err, 1ull); break; }
    }
}

// (flag_base = PB_XCHG_GEN_FLAGS or PB_XCHG_ACK_FLAGS)
__global__ void xchg_flag_signal_kernel(unsigned long long* const* __restrict__ peers, int world, int rank, int flag_base, unsigned long long epoch) {
    const int q = threadIdx.x;
    if (q < world && q != rank) {
        __threadfence_system();
        unsigned long long* f = peers[q] + flag_base + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    }
}
__global__ void xchg_flag_wait_kernel(unsigned long long* const* __restrict__ peers, int world, int rank, int flag_base, unsigned long long epoch) {
    const int q = threadIdx.x;
    if (q < world && q != rank) wait_flag(peers[rank] + flag_base + q, epoch, peers[rank] + PB_XCHG_MAX_WORLD);
}

// grid (x, world): blockIdx.y = the peer whose column slice this block pulls; rows are spread over blockIdx.x
__global__ void __launch_bounds__(256) k3_gather_kernel(unsigned long long* const* __restrict__ peers, uint64_t* const* __restrict__ lab_c,
                                                        uint64_t* const* __restrict__ lab_m, int world, int rank, unsigned long long epoch,
                                                        int64_t buf_off, int32_t P, int64_t Rw, int64_t total_blocks, int64_t wpb) {
    const int q = blockIdx.y;
    if (q == rank) return;
    if (threadIdx.x == 0) wait_flag(peers[rank] + PB_XCHG_GEN_FLAGS + q, epoch, peers[rank] + PB_XCHG_MAX_WORLD);
    __syncthreads();
    const int64_t c0 = (total_blocks * q / world) * wpb, c1 = (total_blocks * (q + 1) / world) * wpb;   // peer q's word columns
    const int64_t pairs = (c1 - c0) >> 1;                       // 16 bytes per thread and step (wpb is even)
    if (pairs <= 0) return;
    const int n_mat = lab_m[q] ? 2 : 1;
    for (int mtx = 0; mtx < n_mat; mtx++) {
        const uint64_t* src = (mtx ? lab_m[q] : lab_c[q]) + buf_off + c0;
        uint64_t* dst = (mtx ? lab_m[rank] : lab_c[rank]) + buf_off + c0;
        // four 16-byte peer loads in flight per thread: an NVLink round trip is microseconds, the link wants megabytes in flight
        const int64_t total = (int64_t)P * pairs, stride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
            unsigned long long a[4], b[4];
            int64_t off[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int64_t i = i0 + k * stride;
                off[k] = -1;
                if (i < total) {
                    const int64_t row = i / pairs, pr = i - row * pairs;
                    off[k] = row * Rw + 2 * pr;
                    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a[k]), "=l"(b[k]) : "l"(src + off[k]) : "memory");
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (off[k] >= 0) *reinterpret_cast<ulonglong2*>(dst + off[k]) = make_ulonglong2(a[k], b[k]);
        }
    }
}

bool pbx_labels_shared(pb_ctx* ctx, size_t need_bytes) {
    const pb_xchg& x = ctx->xchg;
    return x.active && x.lab_active && ctx->cfg.mode == PB_MODE_PHILOX && ctx->d_Xcase == x.lab_c[ctx->rank] &&
           (ctx->n_miss > 0 ? (ctx->d_Xmiss != nullptr && ctx->d_Xmiss == x.lab_m[ctx->rank]) : x.lab_m[ctx->rank] == nullptr) &&
           need_bytes <= x.lab_bytes && !getenv("PB_NO_LABEL_SHARING");
}

int pbx_labels_begin(pb_ctx* ctx, int buf, cudaStream_t st) {
    pb_xchg& x = ctx->xchg;
    if (x.buf_epoch[buf] == 0) return PB_OK;
    xchg_flag_wait_kernel<<<1, PB_XCHG_MAX_WORLD, 0, st>>>(x.d_peer_tbl, ctx->world, ctx->rank, PB_XCHG_ACK_FLAGS, x.buf_epoch[buf]);
    ctx->tm.n_launches += 1;
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pbx_labels_exchange(pb_ctx* ctx, int buf, int64_t buf_off_words, int32_t P, int64_t Rw, int64_t total_blocks, int64_t words_per_block,
                        bool has_miss, cudaStream_t st) {
    pb_xchg& x = ctx->xchg;
    (void)has_miss;
    const unsigned long long e = ++x.gen_epoch;
    xchg_flag_signal_kernel<<<1, PB_XCHG_MAX_WORLD, 0, st>>>(x.d_peer_tbl, ctx->world, ctx->rank, PB_XCHG_GEN_FLAGS, e);
    const int64_t slice_pairs = (total_blocks / ctx->world + 1) * words_per_block / 2;
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(((int64_t)P * slice_pairs + 1023) / 1024, (int64_t)ctx->sm_count));   // x (world - 1) peers in grid.y
    k3_gather_kernel<<<dim3(gx, (unsigned)ctx->world), 256, 0, st>>>(x.d_peer_tbl, x.d_lab_c, x.d_lab_m, ctx->world, ctx->rank, e, buf_off_words, P, Rw,
                                                                   total_blocks, words_per_block);
    xchg_flag_signal_kernel<<<1, PB_XCHG_MAX_WORLD, 0, st>>>(x.d_peer_tbl, ctx->world, ctx->rank, PB_XCHG_ACK_FLAGS, e);
    x.buf_epoch[buf] = e;
    ctx->tm.n_launches += 3;
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

static int labels_finish(pb_ctx* ctx, size_t bytes) {
    pb_xchg& x = ctx->xchg;
    if (!x.d_lab_c) PB_CUDA(cudaMalloc(&x.d_lab_c, PB_XCHG_MAX_WORLD * sizeof(void*)));
    if (!x.d_lab_m) PB_CUDA(cudaMalloc(&x.d_lab_m, PB_XCHG_MAX_WORLD * sizeof(void*)));
    PB_CUDA(cudaMemcpy(x.d_lab_c, x.lab_c, PB_XCHG_MAX_WORLD * sizeof(void*), cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(x.d_lab_m, x.lab_m, PB_XCHG_MAX_WORLD * sizeof(void*), cudaMemcpyHostToDevice));
    x.lab_bytes = bytes;
    x.lab_active = true;
    return PB_OK;
}

static void labels_unmap(pb_ctx* ctx) {
    pb_xchg& x = ctx->xchg;
    for (int q = 0; q < PB_XCHG_MAX_WORLD; q++) {
        if (x.lab_ipc[q]) {
            if (x.lab_c[q]) cudaIpcCloseMemHandle(x.lab_c[q]);
            if (x.lab_m[q]) cudaIpcCloseMemHandle(x.lab_m[q]);
        }
        x.lab_c[q] = x.lab_m[q] = nullptr; x.lab_ipc[q] = false;
    }
    x.lab_active = false;
    cudaGetLastError();
}

// blob: [cudaIpcMemHandle_t of d_Xcase][cudaIpcMemHandle_t of d_Xmiss, zeros without missing labels]
int pbx_labels_export(pb_ctx* ctx, void* blob128) {
    int rc = pbk_labels_alloc(ctx);
    if (rc) return rc;
    if (!ctx->d_Xcase) { ctx->err = "pb_comm_export_labels: no label matrix (Philox mode with labels set is required)"; return PB_ERR_INVALID; }
    memset(blob128, 0, 128);
    cudaIpcMemHandle_t h;
    PB_CUDA(cudaIpcGetMemHandle(&h, ctx->d_Xcase));
    memcpy(blob128, &h, 64);
    if (ctx->d_Xmiss && ctx->n_miss > 0) {
        PB_CUDA(cudaIpcGetMemHandle(&h, ctx->d_Xmiss));
        memcpy((char*)blob128 + 64, &h, 64);
    }
    return PB_OK;
}

int pbx_labels_connect_ipc(pb_ctx* ctx, const void* blobs) {
    pb_xchg& x = ctx->xchg;
    if (!x.active) { ctx->err = "pb_comm_connect_labels: call pb_comm_connect first"; return PB_ERR_INVALID; }
    if (!ctx->d_Xcase) { ctx->err = "pb_comm_connect_labels: call pb_comm_export_labels first"; return PB_ERR_INVALID; }
    labels_unmap(ctx);
    const bool miss = ctx->d_Xmiss && ctx->n_miss > 0;
    for (int q = 0; q < ctx->world; q++) {
        if (q == ctx->rank) { x.lab_c[q] = ctx->d_Xcase; x.lab_m[q] = miss ? ctx->d_Xmiss : nullptr; continue; }
        cudaIpcMemHandle_t h;
        void* p = nullptr;
        memcpy(&h, (const char*)blobs + 128 * (size_t)q, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess && miss) {
            x.lab_c[q] = (uint64_t*)p; x.lab_ipc[q] = true;
            memcpy(&h, (const char*)blobs + 128 * (size_t)q + 64, 64);
            e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e == cudaSuccess) x.lab_m[q] = (uint64_t*)p;
        } else if (e == cudaSuccess) { x.lab_c[q] = (uint64_t*)p; x.lab_ipc[q] = true; }
        if (e != cudaSuccess) {
            ctx->err = std::string("pb_comm_connect_labels: cudaIpcOpenMemHandle (rank ") + std::to_string(q) + "): " + cudaGetErrorString(e);
            cudaGetLastError();
            labels_unmap(ctx);
            return PB_ERR_NCCL;
        }
    }
    return labels_finish(ctx, (size_t)ctx->X_cap);
}

// contexts of ONE process (pb_group), matrices already allocated by pbk_labels_alloc on every one of them
int pbx_labels_connect_local(pb_ctx** ctxs, int n) {
    for (int i = 0; i < n; i++) {
        if (!ctxs[i]->xchg.active || !ctxs[i]->d_Xcase) return PB_OK;          // nothing to share (one device, replay mode, no labels yet)
        for (int j = 0; j < i; j++) if (ctxs[j]->device == ctxs[i]->device) return PB_OK;   // shards on one GPU (test hook): sharing gains nothing
    }
    for (int i = 0; i < n; i++) {
        pb_ctx* ctx = ctxs[i];
        PB_CUDA(cudaSetDevice(ctx->device));
        labels_unmap(ctx);
        const bool miss = ctx->d_Xmiss && ctx->n_miss > 0;
        for (int j = 0; j < n; j++) { ctx->xchg.lab_c[j] = ctxs[j]->d_Xcase; ctx->xchg.lab_m[j] = miss ? ctxs[j]->d_Xmiss : nullptr; }
        size_t bytes = (size_t)ctx->X_cap;
        for (int j = 0; j < n; j++) bytes = std::min(bytes, (size_t)ctxs[j]->X_cap);
        int rc = labels_finish(ctx, bytes);
        if (rc) return rc;
    }
    return PB_OK;
}

int pbx_check(pb_ctx* ctx) {
    pb_xchg& x = ctx->xchg;
    if (!x.active) return PB_OK;
    unsigned long long err = 0;
    PB_CUDA(cudaMemcpy(&err, x.block + PB_XCHG_MAX_WORLD, 8, cudaMemcpyDeviceToHost));
    if (err) {
        PB_CUDA(cudaMemset(x.block + PB_XCHG_MAX_WORLD, 0, 8));
        ctx->err = "pb_run_assoc: a peer rank did not deliver its maxT buffer within 20 s (did every rank call pb_run_assoc?)";
        return PB_ERR_NCCL;
    }
    return PB_OK;
}

void pbx_destroy(pb_ctx* ctx) {
    pb_xchg& x = ctx->xchg;
    labels_unmap(ctx);
    if (x.d_lab_c) cudaFree(x.d_lab_c);
    if (x.d_lab_m) cudaFree(x.d_lab_m);
    x.d_lab_c = x.d_lab_m = nullptr;
    for (int q = 0; q < PB_XCHG_MAX_WORLD; q++)
        if (x.peer_ipc[q] && x.peer[q]) { cudaIpcCloseMemHandle(x.peer[q]); x.peer[q] = nullptr; x.peer_ipc[q] = false; }
    if (x.block) cudaFree(x.block);
    if (x.d_pooled) cudaFree(x.d_pooled);
    if (x.d_peer_tbl) cudaFree(x.d_peer_tbl);
    x.block = nullptr; x.d_pooled = nullptr; x.d_peer_tbl = nullptr;
    x.active = false;
    cudaGetLastError();
}
