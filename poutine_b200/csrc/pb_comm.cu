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
#include <string.h>

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
    for (int q = 0; q < PB_XCHG_MAX_WORLD; q++)
        if (x.peer_ipc[q] && x.peer[q]) { cudaIpcCloseMemHandle(x.peer[q]); x.peer[q] = nullptr; x.peer_ipc[q] = false; }
    if (x.block) cudaFree(x.block);
    if (x.d_pooled) cudaFree(x.d_pooled);
    if (x.d_peer_tbl) cudaFree(x.d_peer_tbl);
    x.block = nullptr; x.d_pooled = nullptr; x.d_peer_tbl = nullptr;
    x.active = false;
    cudaGetLastError();
}
