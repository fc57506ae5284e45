// K1 — homoplasy event detection over node-major bit planes (sm_100a).
//
// Replaces count_all_homoplasy_events / identify_homoplasy_events_for_seg_site
// (HC.java:1402-1511, 1647-1698) and the per-site allele census (get_alleles HC.java:1593-1603,
// identify_major_minor_alleles HC.java:1712-1743).
//
// Closed form of the reference's leaf->root walks: node v is an MRCA iff
//     v != root  &&  parent(v) != root  &&  geno(v) in ACGT  &&  geno(parent(v)) != geno(v)
// plus the root itself; v goes to bucket a1 iff geno(v) == allele1, else a2; a bucket with
// fewer than 2 members is emptied.  The test is edge-local, so K1 streams every plane row ONCE:
//   * a block owns 128 word columns (8,192 sites) and one chunk of the host-built DFS edge program;
//   * a producer warp (one elected lane) moves each row segment (3 planes x 1 KB, one 3-D TMA box) into a shared-memory
//     ring of 8 rows (2 stages of four boxes, cp.async.bulk.tensor + mbarrier complete_tx); the four
//     consumer warps wait on the stage's full barrier, pull their 64-bit words into registers and
//     release the stage at once (empty barrier) — memory-level parallelism comes from the ring,
//     not from registers;
//   * the DFS order keeps the parent row in registers: MRCA mask of an edge =
//     Vc & (~Vp | (B0c^B0p) | (B1c^B1p)) (3 LOP3 per 32-bit half);
//   * events are sparse (~2.5 per site per tree): a (row, word) that holds any goes, masks as they are, into a
//     per-block shared queue of 32-byte records drained with one global atomic per drain; the tail kernels peel the bits;
//   * the leaf census (how many leaves carry A/C/G/T per site) is dense: bit-sliced vertical counters,
//     Harley-Seal carry-save levels 1/2/4 in registers, slices 8.. in shared memory.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "pb_internal.h"

struct K1Params {
    int64_t W;
    const uint32_t* __restrict__ ops;
    const int32_t* __restrict__ chunk_op;
    uint64_t* __restrict__ census;         // [flush][4][KS][W]
    raw_rec_t* __restrict__ raw;
    unsigned long long* raw_cursor;
    unsigned long long raw_cap;
};

static_assert(PB_K1_LANE_BITS == 64, "a consumer thread owns one whole 64-site word column (raw records hold whole words); "
                                     "32-bit lanes were measured slower in round 1 and are no longer supported");
typedef uint64_t lane_t;

// ---- mbarrier / TMA primitives (inline PTX; barrier and ring addresses are 32-bit shared-window offsets) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "K1_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"   // %2: suspend-time hint, the thread sleeps instead of spinning
        "@!p bra K1_WAIT_%=;\n\t}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
// one 3-D tiled TMA load: box (128 words, 1 node, 3 planes) at (col0, node, 0) -> 3 KB in shared memory
__device__ __forceinline__ void tma_load_row(uint32_t dst, const CUtensorMap* tmap, int32_t col0, int32_t node, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tmap), "r"(col0), "r"(node), "r"(0), "r"(bar) : "memory");
}
__device__ __forceinline__ lane_t lds_lane(uint32_t addr) {
    lane_t v;
#if PB_K1_LANE_BITS == 64
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
#else
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
#endif
    return v;
}
// keeps a shared-window address in a register (stops the compiler from re-deriving it from the CTA id every use)
__device__ __forceinline__ uint32_t pin_u32(uint32_t x) {
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

// Queue push: one shared-memory atomic and two 16-byte stores per (row, word) that holds any event; the bits are peeled by
// the tail kernels (k1_count_events_kernel, k1_scatter_events_kernel).
template <int QCAP>
__device__ __forceinline__ void emit_events(lane_t mrca, lane_t b0, lane_t b1, uint32_t site_base, uint32_t node_tag,
                                            uint32_t queue_addr, uint32_t count_addr, const K1Params& p) {
    static_assert(sizeof(lane_t) == 8, "raw records hold whole 64-site words");
    uint32_t slot;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(count_addr) : "memory");
    const uint32_t wcol = site_base >> 6;
    if (slot < QCAP) {
        const uint32_t a = queue_addr + slot * 32;
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(wcol), "r"(node_tag) : "memory");
        asm volatile("st.shared.u64 [%0], %1;" ::"r"(a + 8), "l"(mrca) : "memory");
        asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(a + 16), "l"(b0), "l"(b1) : "memory");
    } else {  // queue full: slow path straight to global (correct for arbitrarily dense inputs)
        const unsigned long long pos = atomicAdd(p.raw_cursor, 1ull);
        if (pos < p.raw_cap) { raw_rec_t r; r.wcol = wcol; r.tag = node_tag; r.mrca = mrca; r.b0 = b0; r.b1 = b1; p.raw[pos] = r; }
    }
}

// consumers only: move the block's queued records (U 16-byte units each) to the global list (one global atomic).
// The drain decision must be block-uniform: every consumer evaluates "queue at least `threshold` full" on its own read of the
// counter BEFORE it arrives, and the arrival barrier ORs the votes (bar.red.or).  All pushes of a consumer precede its own
// arrival, so after the barrier either nobody drains (and consumers may push again at once) or everybody does — then no
// consumer pushes until the last barrier below, and the count every thread reads after the vote is the same, final one.
// (A per-thread re-read of the counter after a plain barrier let a fast consumer push between a slow consumer's barrier
// and its read: the two could disagree about draining and fall out of barrier phase.)
template <int QCAP, int CONSUMERS, int U>
__device__ __noinline__ void drain_queue(const uint4* s_queue, unsigned int* s_count, unsigned long long* s_base, int tid,
                                         unsigned int threshold, void* raw, unsigned long long* raw_cursor, unsigned long long raw_cap) {
    const unsigned int seen = *reinterpret_cast<volatile unsigned int*>(s_count);
    int go;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, 1, %2, q;\n\tselp.s32 %0, 1, 0, p;\n\t}"
                 : "=r"(go) : "r"((unsigned int)(seen >= threshold && seen > 0)), "n"(CONSUMERS) : "memory");
    if (go) {   // block-uniform (the barrier's OR)
        const unsigned int n = min(*reinterpret_cast<volatile unsigned int*>(s_count), (unsigned int)QCAP);
        if (tid == 0) *s_base = atomicAdd(raw_cursor, (unsigned long long)n);
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
        const unsigned long long base = *s_base;
        uint4* out = reinterpret_cast<uint4*>(raw);
        for (unsigned int i = tid; i < U * n; i += CONSUMERS) {       // 16-byte units: coalesced
            const unsigned long long pos = base + i / U;
            if (pos < raw_cap) out[U * base + i] = s_queue[i];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
        if (tid == 0) *s_count = 0;
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
    }
}

#define K1_ROW_BYTES (3 * PB_K1_COLS * 8)
#define K1_STAGE_BYTES (2 * PB_K1_RPS * K1_ROW_BYTES)
#define K1_SMEM_RING (PB_K1_STAGES * K1_STAGE_BYTES)
#define K1_SMEM_HI ((PB_K1_KHI + 1) * 4 * PB_K1_COLS * 8)   // high slices + the pending carries into slice 8
#define K1_SMEM_QUEUE (PB_K1_QCAP * 32)
#define K1_SMEM_BARS (2 * PB_K1_STAGES * 8)
#define K1_SMEM_TOTAL (K1_SMEM_RING + K1_SMEM_HI + K1_SMEM_QUEUE + K1_SMEM_BARS)

// consumer-thread state of the leaf census (see k1_events_kernel)
struct Census {
    lane_t ones[4], twos[4], fours[4], pend2[4], pend4[4];
    int nleaf;
};

// one tree row: MRCA test against the current parent row, optional parent update, optional leaf census
__device__ __forceinline__ void k1_row(uint32_t op, lane_t cv, lane_t c0, lane_t c1, lane_t& pv, lane_t& p0, lane_t& p1,
                                       Census& cs, lane_t* s_hi, lane_t* s_p8, int tid, uint32_t site_base, uint32_t queue_addr,
                                       uint32_t count_addr, const K1Params& p) {
    if (op & PB_OP_TEST) {
        const lane_t mrca = cv & (~pv | (c0 ^ p0) | (c1 ^ p1));
        if (mrca) emit_events<PB_K1_QCAP>(mrca, c0, c1, site_base, (op & PB_OP_NODE_MASK) | (op & PB_OP_LEAF), queue_addr, count_addr, p);
    }
    if (op & PB_OP_SETCUR) { pv = cv; p0 = c0; p1 = c1; }
    if (op & PB_OP_LEAF) {
        // the four counters are n(V), n(V&B0), n(V&B1), n(V&B0&B1) (k1_finalize_kernel turns them into per-base counts):
        // new = old ^ (V & B) and carry = old & ~new are one 3-input LOP3 each, no materialised one-hot masks
        const int ph = cs.nleaf & 7;                                 // block-uniform
        cs.nleaf++;
        {
            const lane_t o0 = cs.ones[0], o1 = cs.ones[1], o2 = cs.ones[2], o3 = cs.ones[3], t3 = cv & c0 & c1;
            cs.ones[0] = o0 ^ cv;        cs.pend2[0] |= o0 & ~cs.ones[0];
            cs.ones[1] = o1 ^ (cv & c0); cs.pend2[1] |= o1 & ~cs.ones[1];
            cs.ones[2] = o2 ^ (cv & c1); cs.pend2[2] |= o2 & ~cs.ones[2];
            cs.ones[3] = o3 ^ t3;        cs.pend2[3] |= o3 & t3;
        }
        if (ph & 1) {
#pragma unroll
            for (int c = 0; c < 4; c++) { cs.pend4[c] |= cs.twos[c] & cs.pend2[c]; cs.twos[c] ^= cs.pend2[c]; cs.pend2[c] = 0; }
            if (ph & 2) {
                if (!(ph & 4)) {
#pragma unroll
                    for (int c = 0; c < 4; c++) { s_p8[c * PB_K1_CONSUMERS] = cs.fours[c] & cs.pend4[c]; cs.fours[c] ^= cs.pend4[c]; cs.pend4[c] = 0; }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        lane_t carry = s_p8[c * PB_K1_CONSUMERS] | (cs.fours[c] & cs.pend4[c]);
                        cs.fours[c] ^= cs.pend4[c]; cs.pend4[c] = 0;
#pragma unroll
                        for (int j = 0; j < PB_K1_KHI; j++) {
                            lane_t* hp = s_hi + (j * 4 + c) * PB_K1_CONSUMERS + tid;
                            const lane_t h = *hp;
                            *hp = h ^ carry;
                            carry &= h;
                        }
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(PB_K1_THREADS, PB_K1_BLOCKS_PER_SM) k1_events_kernel(const __grid_constant__ CUtensorMap tmap, const K1Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    lane_t* s_hi = reinterpret_cast<lane_t*>(smem + K1_SMEM_RING);                        // [slice][base][consumer]
    const uint4* s_queue = reinterpret_cast<const uint4*>(smem + K1_SMEM_RING + K1_SMEM_HI);
    __shared__ unsigned int s_count;
    __shared__ unsigned long long s_base;

    const int tid = threadIdx.x;
    const int chunk = blockIdx.y;
    const int32_t col0 = (int32_t)blockIdx.x * PB_K1_COLS;
    const int r0 = p.chunk_op[chunk], r1 = p.chunk_op[chunk + 1];
    const uint4* __restrict__ recs = reinterpret_cast<const uint4*>(p.ops);
    const uint32_t ring0 = pin_u32(smem_u32(smem));
    const uint32_t full0 = pin_u32(ring0 + K1_SMEM_RING + K1_SMEM_HI + K1_SMEM_QUEUE), empty0 = full0 + 8 * PB_K1_STAGES;

    if (tid == 0) {
        s_count = 0;
        for (int s = 0; s < PB_K1_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, PB_K1_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < PB_K1_CONSUMERS)
        for (int i = 0; i < (PB_K1_KHI + 1) * 4; i++) s_hi[i * PB_K1_CONSUMERS + tid] = 0;
    __syncthreads();

    if (tid >= PB_K1_CONSUMERS) {
        // ================= producer warp: one lane feeds the ring with TMA =================
        if (tid == PB_K1_CONSUMERS) {
            uint32_t soff = 0, boff = 0, phase = 0;
            for (int i = r0; i < r1; i += PB_K1_RPS) {
                // padding slots carry node = PB_OP_NULL, a coordinate beyond the tensor's node extent: TMA zero-fills the box
                // without touching memory and still counts its bytes, so every stage is four loads and a constant tx count
                uint32_t node[2 * PB_K1_RPS];
#pragma unroll
                for (int k = 0; k < PB_K1_RPS; k++) {
                    const uint4 rec = __ldg(recs + i + k);
                    node[2 * k] = rec.x & PB_OP_NODE_MASK; node[2 * k + 1] = rec.y & PB_OP_NODE_MASK;
                }
                mbar_wait(empty0 + boff, phase ^ 1);                     // consumers have drained this stage
                mbar_arrive_expect_tx(full0 + boff, K1_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < 2 * PB_K1_RPS; k++) tma_load_row(ring0 + soff + k * K1_ROW_BYTES, &tmap, col0, (int32_t)node[k], full0 + boff);
                soff += K1_STAGE_BYTES; boff += 8;
                if (boff == 8 * PB_K1_STAGES) { soff = 0; boff = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ================= consumer warps =================
    const int lane = tid & 31;
    constexpr int SUBS = 64 / PB_K1_LANE_BITS;                          // consumer threads per word column
    const int sub = tid % SUBS;                                         // which part of the 64-site word this thread owns
    const int64_t w = (int64_t)col0 + tid / SUBS;
    const uint32_t site_base = (uint32_t)(w * 64 + sub * PB_K1_LANE_BITS);
    const uint32_t my_ring = pin_u32(ring0 + tid * (PB_K1_LANE_BITS / 8));
    const uint32_t queue_addr = pin_u32(ring0 + K1_SMEM_RING + K1_SMEM_HI), count_addr = pin_u32(smem_u32(&s_count));

    // leaf census, bit-sliced with deferred carries: level L holds a counter slice and the OR of the carries the
    // level below produced since its last update (two consecutive half-add carries are mutually exclusive:
    // after a carry the slice bit is 0), and is updated every 2^L leaves.  No buffers to copy, 2 LOP3 per level-1 add.
    Census cs;
#pragma unroll
    for (int c = 0; c < 4; c++) { cs.ones[c] = cs.twos[c] = cs.fours[c] = 0; cs.pend2[c] = cs.pend4[c] = 0; }
    cs.nleaf = 0;
    lane_t* s_p8 = s_hi + (PB_K1_KHI * 4) * PB_K1_CONSUMERS + tid;    // pending carries into slice 8
    lane_t pv = 0, p0 = 0, p1 = 0;   // current parent row
    uint32_t soff = 0, boff = 0, phase = 0;

    // Rows arrive clean: padding slots are zero-filled by TMA (no allele, no event), columns beyond W likewise, and the
    // caller's padding bits of word W-1 were cleared from the V plane at upload (pb_put_planes -> pbk_mask_last_word).
#pragma unroll 1
    for (int i = r0; i < r1; i += PB_K1_RPS) {
        uint4 rec[PB_K1_RPS];
#pragma unroll
        for (int k = 0; k < PB_K1_RPS; k++) rec[k] = __ldg(recs + i + k);
        mbar_wait(full0 + boff, phase);                                  // the stage's TMA bytes have landed
        const uint32_t src = my_ring + soff;
#pragma unroll
        for (int k = 0; k < PB_K1_RPS; k++) {
            const uint32_t s2 = src + k * 2 * K1_ROW_BYTES;
            const lane_t av = lds_lane(s2), a0 = lds_lane(s2 + PB_K1_COLS * 8), a1 = lds_lane(s2 + 2 * PB_K1_COLS * 8);
            const lane_t bv = lds_lane(s2 + 3 * PB_K1_COLS * 8), b0 = lds_lane(s2 + 4 * PB_K1_COLS * 8), b1 = lds_lane(s2 + 5 * PB_K1_COLS * 8);
            if (k == PB_K1_RPS - 1) {                                    // every row of the stage is in registers: free it
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + boff);
                soff += K1_STAGE_BYTES; boff += 8;
                if (boff == 8 * PB_K1_STAGES) { soff = 0; boff = 0; phase ^= 1; }
            }
            k1_row(rec[k].x, av, a0, a1, pv, p0, p1, cs, s_hi, s_p8, tid, site_base, queue_addr, count_addr, p);
            k1_row(rec[k].y, bv, b0, b1, pv, p0, p1, cs, s_hi, s_p8, tid, site_base, queue_addr, count_addr, p);
            if (rec[k].z != PB_REC_NOFLUSH) {
                // census partial rec.z: [flush][base][slice][W]   (segments are padded to 8 leaves: nothing is pending)
                if (w < p.W) {
                    // census words are uint64 [flush][base][slice][W]; a 32-bit lane writes its half (little-endian)
                    lane_t* out = reinterpret_cast<lane_t*>(p.census) + (((int64_t)rec[k].z * 4 * PB_K1_KS) * p.W + w) * SUBS + sub;
                    const int64_t sl = p.W * SUBS;                       // lanes per slice
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        out[((int64_t)c * PB_K1_KS + 0) * sl] = cs.ones[c];
                        out[((int64_t)c * PB_K1_KS + 1) * sl] = cs.twos[c];
                        out[((int64_t)c * PB_K1_KS + 2) * sl] = cs.fours[c];
#pragma unroll
                        for (int j = 0; j < PB_K1_KHI; j++) out[((int64_t)c * PB_K1_KS + 3 + j) * sl] = s_hi[(j * 4 + c) * PB_K1_CONSUMERS + tid];
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; c++) { cs.ones[c] = cs.twos[c] = cs.fours[c] = 0; cs.pend2[c] = cs.pend4[c] = 0; }
#pragma unroll
                for (int j = 0; j < (PB_K1_KHI + 1) * 4; j++) s_hi[j * PB_K1_CONSUMERS + tid] = 0;
                cs.nleaf = 0;
            }
        }
        if ((((i - r0) / PB_K1_RPS) & (PB_K1_DRAIN_EVERY / PB_K1_RPS - 1)) == PB_K1_DRAIN_EVERY / PB_K1_RPS - 1)
            drain_queue<PB_K1_QCAP, PB_K1_CONSUMERS, 2>(s_queue, &s_count, &s_base, tid, PB_K1_QCAP / 2, p.raw, p.raw_cursor, p.raw_cap);
    }
    drain_queue<PB_K1_QCAP, PB_K1_CONSUMERS, 2>(s_queue, &s_count, &s_base, tid, 1, p.raw, p.raw_cursor, p.raw_cap);
}

// =================================================================================================
// K1X — the sweep over ONE plane.  When every site carries at most two bases and every genotype is valid
// (pb_put_planes_biallelic with V = NULL for every tile), a node's genotype is one bit: X = "carries the site's second
// base", a tree row is a third of the bytes and the MRCA test of an edge is X_child ^ X_parent.  With so little work per
// word the per-row control flow (op decode, ring handshake) dominates, so
//   * a consumer thread owns PB_KX_WPT adjacent word columns (one 16-byte shared-memory load per row) and a block
//     PB_KX_COLS = 128 * PB_KX_WPT columns; its own lowering of the edge list (pb_ctx::d_ops_x) is cut for that grid;
//   * there is NO leaf census in the sweep.  X(leaf) = X(root) + the sum of X(c) - X(p) over the edges of its root path, so
//       #leaves carrying the second base = L * X(root) + sum over edges with X(c) != X(p) of (X(c) ? +1 : -1) * leaves_below(c):
//     a sum over the sparse records the sweep emits anyway (k1_count_events_kernel adds it up, one atomic per event).  For that
//     the edges hanging off the root are tested too; their records are marked NOMRCA (the reference never makes a child of the
//     root an MRCA, HC.java:1674) and feed the census only.  The dense bit-sliced counter this replaces was a fifth of the
//     sweep's instructions and 20 of its registers;
//   * a record is 16 bytes {word column, node | X | flags, site mask} (xrec_t): one shared atomic + ONE 16-byte store; a word
//     whose edge flips sites in both directions gives two records.
// Same ring protocol as K1 (2*RPS TMA boxes per stage, PB_KX_STAGES stages, here 2-D boxes), same queue drain.
#ifndef PB_KX_WPT
#define PB_KX_WPT 2
#endif
#ifndef PB_KX_STAGES
#define PB_KX_STAGES (4 / PB_KX_WPT)
#endif
#ifndef PB_KX_BLOCKS_PER_SM
#define PB_KX_BLOCKS_PER_SM 7
#endif
#ifndef PB_KX_DRAIN_EVERY
#define PB_KX_DRAIN_EVERY PB_K1_DRAIN_EVERY
#endif
#define PB_KX_COLS (PB_K1_COLS * PB_KX_WPT)
#define KX_ROW_BYTES (PB_KX_COLS * 8)
#define KX_STAGE_BYTES (2 * PB_K1_RPS * KX_ROW_BYTES)
#define KX_SMEM_RING (PB_KX_STAGES * KX_STAGE_BYTES)
#define KX_SMEM_BARS (2 * PB_KX_STAGES * 8)
#define PB_KX_QCAP (PB_K1_QCAP * PB_KX_WPT + 128)              // record queue: a block sees PB_KX_WPT times the events per row
#define KX_SMEM_QUEUE (PB_KX_QCAP * 16)
#define KX_SMEM_TOTAL (KX_SMEM_RING + KX_SMEM_QUEUE + KX_SMEM_BARS)

struct KXParams {
    int32_t col_block0;                   // first column block of this launch (tiles swept ahead of pb_run_events cover a sub-range)
    const uint32_t* __restrict__ ops;
    const int32_t* __restrict__ chunk_op;
    xrec_t* __restrict__ raw;
    unsigned long long* raw_cursor;
    unsigned long long raw_cap;
};

__device__ __forceinline__ void tma_load_row_2d(uint32_t dst, const CUtensorMap* tmap, int32_t col0, int32_t node, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(col0), "r"(node), "r"(bar) : "memory");
}

struct XRow { uint64_t w[PB_KX_WPT]; };

__device__ __forceinline__ XRow lds_xrow(uint32_t addr) {
    XRow r;
#if PB_KX_WPT == 2
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(r.w[0]), "=l"(r.w[1]) : "r"(addr));
#else
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r.w[0]) : "r"(addr));
#endif
    return r;
}

// queue full: slow path straight to global (correct for arbitrarily dense inputs); out of line, it is rare
__device__ __noinline__ void emit_xrec_overflow(uint64_t mask, uint32_t wcol, uint32_t tag, xrec_t* raw, unsigned long long* raw_cursor,
                                                unsigned long long raw_cap) {
    const unsigned long long pos = atomicAdd(raw_cursor, 1ull);
    if (pos < raw_cap) { xrec_t r; r.wcol = wcol; r.tag = tag; r.mask = mask; raw[pos] = r; }
}

// queue push of one xrec_t: one shared-memory atomic and one 16-byte store
#ifdef PB_KX_EMIT_NOINLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
void emit_xrec(uint64_t mask, uint32_t wcol, uint32_t tag, uint32_t queue_addr, uint32_t count_addr, const KXParams& q) {
    uint32_t slot;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(count_addr) : "memory");
    if (__builtin_expect(slot < PB_KX_QCAP, 1)) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(queue_addr + slot * 16), "r"(wcol), "r"(tag), "r"((uint32_t)mask),
                     "r"((uint32_t)(mask >> 32)) : "memory");
    } else {
        emit_xrec_overflow(mask, wcol, tag, q.raw, q.raw_cursor, q.raw_cap);
    }
}

// one tree row: X_child ^ X_parent against the current parent row, optional parent update
__device__ __forceinline__ void k1x_row(uint32_t op, const XRow& x, XRow& px, uint32_t wcol0, uint32_t queue_addr, uint32_t count_addr,
                                        const KXParams& q) {
    if (op & PB_OPX_TEST) {
        const uint32_t tag = op & PB_OPX_TAG_MASK;                 // node | NOMRCA | LEAF sit where xrec_t wants them; bit 28 = the child's X
#pragma unroll
        for (int j = 0; j < PB_KX_WPT; j++) {
            const uint64_t diff = x.w[j] ^ px.w[j];
            if (diff) {
                const uint64_t up = diff & x.w[j];                 // sites where the child gained the second base
                if (up) emit_xrec(up, wcol0 + j, tag | PB_XT_XBIT, queue_addr, count_addr, q);
                if (diff != up) emit_xrec(diff ^ up, wcol0 + j, tag, queue_addr, count_addr, q);
            }
        }
    }
    if (op & PB_OPX_SETCUR) px = x;
}

__global__ void __launch_bounds__(PB_K1_THREADS, PB_KX_BLOCKS_PER_SM) k1x_events_kernel(const __grid_constant__ CUtensorMap tmap, const KXParams q) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned int s_count;
    __shared__ unsigned long long s_base;
    const uint4* s_queue = reinterpret_cast<const uint4*>(smem + KX_SMEM_RING);

    const int tid = threadIdx.x;
    const int chunk = blockIdx.y;
    const int32_t col0 = ((int32_t)blockIdx.x + q.col_block0) * PB_KX_COLS;
    const int r0 = q.chunk_op[chunk], r1 = q.chunk_op[chunk + 1];
    const uint4* __restrict__ recs = reinterpret_cast<const uint4*>(q.ops);
    const uint32_t ring0 = pin_u32(smem_u32(smem));
    const uint32_t full0 = pin_u32(ring0 + KX_SMEM_RING + KX_SMEM_QUEUE), empty0 = full0 + 8 * PB_KX_STAGES;

    if (tid == 0) {
        s_count = 0;
        for (int s = 0; s < PB_KX_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, PB_K1_COLS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (tid >= PB_K1_COLS) {
        // producer warp: one lane, 2*RPS row boxes per stage (padding slots are out-of-bounds boxes: zero-filled)
        if (tid == PB_K1_COLS) {
            uint32_t soff = 0, boff = 0, phase = 0;
            for (int i = r0; i < r1; i += PB_K1_RPS) {
                uint32_t node[2 * PB_K1_RPS];
#pragma unroll
                for (int k = 0; k < PB_K1_RPS; k++) {
                    const uint4 rec = __ldg(recs + i + k);
                    node[2 * k] = rec.x & PB_OPX_NODE_MASK; node[2 * k + 1] = rec.y & PB_OPX_NODE_MASK;
                }
                mbar_wait(empty0 + boff, phase ^ 1);
                mbar_arrive_expect_tx(full0 + boff, KX_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < 2 * PB_K1_RPS; k++) tma_load_row_2d(ring0 + soff + k * KX_ROW_BYTES, &tmap, col0, (int32_t)node[k], full0 + boff);
                soff += KX_STAGE_BYTES; boff += 8;
                if (boff == 8 * PB_KX_STAGES) { soff = 0; boff = 0; phase ^= 1; }
            }
        }
        return;
    }

    // consumer warps: one thread = PB_KX_WPT adjacent 64-site word columns
    const int lane = tid & 31;
    const uint32_t wcol0 = (uint32_t)col0 + (uint32_t)tid * PB_KX_WPT;
    const uint32_t my_ring = pin_u32(ring0 + tid * (8 * PB_KX_WPT));
    const uint32_t queue_addr = pin_u32(ring0 + KX_SMEM_RING), count_addr = pin_u32(smem_u32(&s_count));

    XRow px;
#pragma unroll
    for (int j = 0; j < PB_KX_WPT; j++) px.w[j] = 0;
    uint32_t soff = 0, boff = 0, phase = 0;

#pragma unroll 1
    for (int i = r0; i < r1; i += PB_K1_RPS) {
        uint4 rec[PB_K1_RPS];
#pragma unroll
        for (int k = 0; k < PB_K1_RPS; k++) rec[k] = __ldg(recs + i + k);
        mbar_wait(full0 + boff, phase);
        const uint32_t src = my_ring + soff;
        XRow x[2 * PB_K1_RPS];
#pragma unroll
        for (int k = 0; k < 2 * PB_K1_RPS; k++) x[k] = lds_xrow(src + k * KX_ROW_BYTES);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + boff);                     // the stage's rows are in registers: free it
        soff += KX_STAGE_BYTES; boff += 8;
        if (boff == 8 * PB_KX_STAGES) { soff = 0; boff = 0; phase ^= 1; }
#pragma unroll
        for (int k = 0; k < PB_K1_RPS; k++) {
            k1x_row(rec[k].x, x[2 * k], px, wcol0, queue_addr, count_addr, q);
            k1x_row(rec[k].y, x[2 * k + 1], px, wcol0, queue_addr, count_addr, q);
        }
        if ((((i - r0) / PB_K1_RPS) & (PB_KX_DRAIN_EVERY / PB_K1_RPS - 1)) == PB_KX_DRAIN_EVERY / PB_K1_RPS - 1)
            drain_queue<PB_KX_QCAP, PB_K1_COLS, 1>(s_queue, &s_count, &s_base, tid, PB_KX_QCAP / 2, q.raw, q.raw_cursor, q.raw_cap);
    }
    drain_queue<PB_KX_QCAP, PB_K1_COLS, 1>(s_queue, &s_count, &s_base, tid, 1, q.raw, q.raw_cursor, q.raw_cap);
}

int pbk_kx_cols() { return PB_KX_COLS; }
int pbk_kx_blocks_per_sm() { return PB_KX_BLOCKS_PER_SM; }

// ---- per-site event counters -------------------------------------------------------------------
// The tail kernels peel the records: one thread per record, one iteration per set bit (usually one).  Both record
// layouts (raw_rec_t of the three-plane sweep, xrec_t of the one-plane sweep) are read into one view.
struct ev_rec_t {
    uint32_t wcol, node;
    bool leaf, nomrca;    // nomrca: one-plane record of an edge hanging off the root (census only)
    uint32_t xbit;        // one-plane mode: the child's X at every site of the mask
    uint64_t mask, b0, b1;
};

template <bool XMODE>
__device__ __forceinline__ ev_rec_t load_rec(const raw_rec_t* __restrict__ raw, int64_t i) {
    ev_rec_t e;
    if (XMODE) {
        const xrec_t r = reinterpret_cast<const xrec_t*>(raw)[i];
        e.wcol = r.wcol; e.node = r.tag & PB_XT_NODE_MASK; e.leaf = (r.tag & PB_XT_LEAF) != 0; e.nomrca = (r.tag & PB_XT_NOMRCA) != 0;
        e.xbit = (r.tag & PB_XT_XBIT) ? 1u : 0u; e.mask = r.mask; e.b0 = e.b1 = 0;
    } else {
        const raw_rec_t r = raw[i];
        e.wcol = r.wcol; e.node = r.tag & PB_EV_NODE_MASK; e.leaf = (r.tag & 0x80000000u) != 0; e.nomrca = false;
        e.xbit = 0; e.mask = r.mrca; e.b0 = r.b0; e.b1 = r.b1;
    }
    return e;
}

// 2-bit base code of the record's node at site bit b; one-plane mode: the site's first base, XOR (first ^ second) if X
template <bool XMODE>
__device__ __forceinline__ uint32_t event_base(const ev_rec_t& e, int b, const uint64_t* __restrict__ xmask) {
    if (!XMODE) return (uint32_t)((e.b0 >> b) & 1ull) | ((uint32_t)((e.b1 >> b) & 1ull) << 1);
    const uint64_t* m = xmask + 4 * (int64_t)e.wcol;
    const uint32_t a = (uint32_t)((m[0] >> b) & 1ull) | ((uint32_t)((m[1] >> b) & 1ull) << 1);
    const uint32_t d = (uint32_t)((m[2] >> b) & 1ull) | ((uint32_t)((m[3] >> b) & 1ull) << 1);
    return a ^ (e.xbit ? d : 0u);
}

// the record list's length lives on the device (the sweep's cursor, clamped to the buffer): no host round trip before the tail.
// One-plane mode also accumulates the census sum (see xrec_t): xcensus[site] += (X_child ? +1 : -1) * leaves_below(child).
template <bool XMODE>
__global__ void k1_count_events_kernel(const raw_rec_t* __restrict__ raw, const unsigned long long* __restrict__ begin,
                                       const unsigned long long* __restrict__ cursor, int64_t raw_cap,
                                       const int32_t* __restrict__ node_sample, const int32_t* __restrict__ leaves_below,
                                       const uint64_t* __restrict__ xmask, uint32_t* __restrict__ cnt, int32_t* __restrict__ xcensus,
                                       unsigned long long* __restrict__ n_events) {
    // records [*begin, *cursor): the whole list, or what one uploaded tile's sweep appended (tail run underneath the upload)
    const int64_t n = min((int64_t)*cursor, raw_cap), i0 = min((int64_t)*begin, n);
    unsigned int nev = 0;
    for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const ev_rec_t r = load_rec<XMODE>(raw, i);
        const bool sampled = r.leaf && node_sample[r.node] >= 0;
        int32_t delta = 0;
        if (XMODE) { const int32_t lb = leaves_below[r.node]; delta = r.xbit ? lb : -lb; }
        uint64_t m = r.mask;
        if (!r.nomrca) nev += (unsigned int)__popcll(m);
        while (m) {
            const int b = __ffsll((long long)m) - 1;
            m &= m - 1;
            const int64_t site = (int64_t)r.wcol * 64 + b;
            if (XMODE) atomicAdd(xcensus + site, delta);
            if (r.nomrca) continue;
            const uint32_t base = event_base<XMODE>(r, b, xmask);
            uint32_t* c = cnt + site * PB_CNT_STRIDE;
            atomicAdd(c + base, 1u);
            if (r.leaf) {
                atomicAdd(c + 4 + base, 1u);
                if (sampled) atomicAdd(c + 8 + base, 1u);
            }
        }
    }
    // one global atomic per block (all blocks hit the same word)
    __shared__ unsigned int s_nev;
    if (threadIdx.x == 0) s_nev = 0;
    __syncthreads();
    nev = __reduce_add_sync(0xffffffffu, nev);
    if ((threadIdx.x & 31) == 0 && nev) atomicAdd(&s_nev, nev);
    __syncthreads();
    if (threadIdx.x == 0 && s_nev) atomicAdd(n_events, (unsigned long long)s_nev);
}

// ---- census reduction: sum the bit-sliced partials of all flushes -----------------------------------
// thread = (word column, base); total[base][PB_K1_KT][W]
__global__ void k1_census_reduce_kernel(const uint64_t* __restrict__ census, int n_flush, int64_t W, uint64_t* __restrict__ total) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y, n_counters = gridDim.y;   // 4 (k1_events_kernel) or 1 (k1x_events_kernel)
    if (w >= W) return;
    uint64_t acc[PB_K1_KT];
#pragma unroll
    for (int j = 0; j < PB_K1_KT; j++) acc[j] = 0;
    for (int f = 0; f < n_flush; f++) {
        const uint64_t* in = census + (((int64_t)f * n_counters + c) * PB_K1_KS) * W + w;
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < PB_K1_KS; j++) {
            const uint64_t b = in[(int64_t)j * W];
            const uint64_t a = acc[j], u = a ^ b;
            acc[j] = u ^ carry;
            carry = (a & b) | (u & carry);
        }
#pragma unroll
        for (int j = PB_K1_KS; j < PB_K1_KT; j++) {
            const uint64_t a = acc[j];
            acc[j] = a ^ carry;
            carry &= a;
        }
    }
    uint64_t* out = total + ((int64_t)c * PB_K1_KT) * W + w;
#pragma unroll
    for (int j = 0; j < PB_K1_KT; j++) out[(int64_t)j * W] = acc[j];
}

// ---- per-site finalisation + informative-site compaction, ONE pass (decoupled look-back scan) ----------------
// census -> class / major-minor (get_alleles HC.java:1593-1603, identify_major_minor_alleles HC.java:1712-1743); counters ->
// Homoplasy_Events (HC.java:5558-5600); subset_by_min_hcount (HC.java:3783-3905) as a stream compaction that keeps site order.
// One thread = one site, one block = one tile of TAIL_TILE sites.  The exclusive prefix (informative sites before this one,
// extant events before this one) is a single-pass scan: every tile publishes its aggregate, then looks back over its
// predecessors' aggregates / inclusive prefixes (Merrill & Garland), so finalize + 3 scan kernels + the site scatter of
// round 1 are one launch.  Tile ids come from an atomic ticket: a tile only waits for tiles that already run.
// XMODE: the alignment is held as one plane.  B0 = X, B1 = the site masks [W][4], V = the valid-site words [W]; the
// census is xcensus[S], the signed sum k1_count_events_kernel took over the records (see xrec_t): leaves carrying the
// second base = n_leaves * X(root) + xcensus.
#define TAIL_TILE 512
#define TS_VALID (1ull << 63)
// scan state: word 0 = ticket, then per tile {aggregate: valid | inf << 40 | ev, prefix inf: valid | n, prefix ev: valid | n}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

template <bool XMODE>
__global__ void __launch_bounds__(TAIL_TILE) k1_finalize_compact_kernel(
    const uint64_t* __restrict__ total, const int32_t* __restrict__ xcensus, int kt_used, int64_t W, int64_t S, const uint64_t* __restrict__ B0,
    const uint64_t* __restrict__ B1, const uint64_t* __restrict__ V, int64_t Wp, int32_t root, int32_t n_leaves, const uint32_t* __restrict__ cnt,
    int32_t min_hcount, int64_t site_offset, pb_site_out* __restrict__ sites, uint8_t* __restrict__ flags, uint64_t* __restrict__ scan_out,
    int32_t* __restrict__ inf_site, AlleleMeta* __restrict__ alleles, unsigned long long* __restrict__ state,
    unsigned long long* __restrict__ class_counts /* [0..4] classes, [5] n_max, [6] A_inuse, [7] S_inf << 38 | E */) {
    __shared__ unsigned int s_cls[5], s_nmax, s_ause;
    __shared__ unsigned long long s_warp[TAIL_TILE / 32], s_excl;
    __shared__ unsigned int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 5) s_cls[tid] = 0;
    if (tid == 0) { s_nmax = 0; s_ause = 0; s_tile = (unsigned int)atomicAdd(state, 1ull); }
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t s = tile * TAIL_TILE + tid;
    uint64_t mine = 0;          // inf << 40 | extant events of the in-use alleles
    uint8_t fl = 0;
    uint32_t n1 = 0, n2 = 0;
    bool u1 = false, u2 = false;
    if (s < S) {
        const int64_t w = s >> 6;
        const int b = (int)(s & 63);
        uint32_t leafcnt[4] = {0, 0, 0, 0};  // by 2-bit code: A C G T
        int xa = 0, xd = 0;                  // XMODE: code of the site's first base, first XOR second
        bool xvalid = true;
        if (XMODE) {
            const uint32_t nx = (uint32_t)((((B0[(int64_t)root * Wp + w] >> b) & 1ull) ? n_leaves : 0) + xcensus[s]);
            const uint64_t* m = B1 + 4 * w;
            xa = (int)((m[0] >> b) & 1ull) | ((int)((m[1] >> b) & 1ull) << 1);
            xd = (int)((m[2] >> b) & 1ull) | ((int)((m[3] >> b) & 1ull) << 1);
            xvalid = (V[w] >> b) & 1ull;     // the column was uploaded (nothing uploaded = no valid genotype, as in full mode)
            if (xvalid) { leafcnt[xa] = (uint32_t)n_leaves - nx; leafcnt[xa ^ xd] += nx; }   // xd == 0 implies nx == 0
        } else {
            uint32_t q[4] = {0, 0, 0, 0};    // K1's counters: leaves with V, V&B0, V&B1, V&B0&B1
#pragma unroll
            for (int c = 0; c < 4; c++)
                for (int j = 0; j < kt_used; j++)
                    q[c] |= (uint32_t)((total[((int64_t)c * PB_K1_KT + j) * W + w] >> b) & 1ull) << j;
            leafcnt[0] = q[0] - q[1] - q[2] + q[3]; leafcnt[1] = q[1] - q[3]; leafcnt[2] = q[2] - q[3]; leafcnt[3] = q[3];
        }
        // HashMap<Character,..> iteration order A, C, T, G  (buckets 'A'&15=1, 'C'&15=3, 'T'&15=4, 'G'&15=7)
        const int order[4] = {0, 1, 3, 2};
        int n_alleles = 0, first = -1, second = -1;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (leafcnt[order[i]] > 0) {
                if (n_alleles == 0) first = order[i];
                else if (n_alleles == 1) second = order[i];
                n_alleles++;
            }
        pb_site_out o;
        o.segsite_id = (int32_t)(site_offset + s + 1);
        o.site_class = n_alleles;
        o.allele1 = o.allele2 = o.a1_count = o.a2_count = o.a1_extant = o.a2_extant = 0;
        if (n_alleles == 2) {
            // identify_major_minor_alleles: first-in-order wins ties (>=, HC.java:1725)
            int a1, a2;
            if (leafcnt[first] >= leafcnt[second]) { a1 = first; a2 = second; } else { a1 = second; a2 = first; }
            const char L4[4] = {'A', 'C', 'G', 'T'};
            const uint32_t* c = cnt + s * PB_CNT_STRIDE;
            const int64_t ro = (int64_t)root * Wp + w;
            const bool rvalid = XMODE ? xvalid : (bool)((V[ro] >> b) & 1ull);
            const int rcode = XMODE ? (xa ^ (((B0[ro] >> b) & 1ull) ? xd : 0))
                                    : ((int)((B0[ro] >> b) & 1ull) | ((int)((B1[ro] >> b) & 1ull) << 1));
            const bool root_a1 = rvalid && rcode == a1;   // add_MRCA(root): a1 iff geno(root)==allele1 (HC.java:1689, 1861)
            uint32_t m_all = c[0] + c[1] + c[2] + c[3], x_all = c[4] + c[5] + c[6] + c[7], xs_all = c[8] + c[9] + c[10] + c[11];
            uint32_t a1c = c[a1] + (root_a1 ? 1u : 0u);
            uint32_t a2c = (m_all - c[a1]) + (root_a1 ? 0u : 1u);
            uint32_t a1x = c[4 + a1], a2x = x_all - c[4 + a1];
            uint32_t a1s = c[8 + a1], a2s = xs_all - c[8 + a1];
            if (a1c < 2) { a1c = 0; a1x = 0; a1s = 0; }   // remove_non_homoplasic_mutations (HC.java:1954-1962)
            if (a2c < 2) { a2c = 0; a2x = 0; a2s = 0; }
            o.allele1 = L4[a1]; o.allele2 = L4[a2];
            o.a1_count = (int32_t)a1c; o.a2_count = (int32_t)a2c; o.a1_extant = (int32_t)a1x; o.a2_extant = (int32_t)a2x;
            // subset_by_min_hcount with EXTANT_NODES_ONLY (HC.java:3783-3905)
            u1 = (int32_t)a1x >= min_hcount; u2 = (int32_t)a2x >= min_hcount;
            // an in-use allele keeps all its sampled leaf events unless its bucket was emptied by the <2 rule (possible only with
            // min_hcount == 0, where in_use does not imply a surviving bucket): a1s / a2s are already 0 then
            n1 = u1 ? a1s : 0u; n2 = u2 ? a2s : 0u;
            fl = PB_SF_BI | (u1 ? PB_SF_A1_USE : 0) | (u2 ? PB_SF_A2_USE : 0) | (uint8_t)(a1 << PB_SF_A1_CODE_SHIFT) |
                 (a1c >= 2 ? PB_SF_A1_ALIVE : 0) | (a2c >= 2 ? PB_SF_A2_ALIVE : 0) | ((u1 || u2) ? PB_SF_INF : 0);
            if (u1 || u2) mine = (1ull << 40) | (uint64_t)(n1 + n2);
        }
        sites[s] = o;
        flags[s] = fl;
        atomicAdd(&s_cls[n_alleles], 1u);
    }
    // ---- block-wide exclusive scan of `mine` (per-tile sums fit: inf <= 512, events < 2^40)
    uint64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t v = lane < TAIL_TILE / 32 ? s_warp[lane] : 0, iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        if (lane < TAIL_TILE / 32) s_warp[lane] = iv - v;     // exclusive warp offsets
        const uint64_t agg = __shfl_sync(0xffffffffu, iv, 31);   // the tile's aggregate
        // ---- decoupled look-back (warp 0)
        unsigned long long* st = state + 1;
        uint64_t excl_inf = 0, excl_ev = 0;
        if (tile > 0) {
            if (lane == 0) { *reinterpret_cast<volatile unsigned long long*>(st + 3 * tile) = TS_VALID | agg; }
            int64_t idx = tile - 1;
            for (;;) {
                const int64_t t = idx - lane;
                uint64_t ag = 0, pi = 0, pe = 0;
                bool has_p = true, ready = true;   // lanes beyond tile 0 count as a ready zero prefix
                if (t >= 0) {
                    do {
                        pi = ld_volatile_u64(st + 3 * t + 1); pe = ld_volatile_u64(st + 3 * t + 2);
                        has_p = (pi & TS_VALID) && (pe & TS_VALID);
                        ag = has_p ? 0 : ld_volatile_u64(st + 3 * t);
                        ready = has_p || (ag & TS_VALID);
                    } while (!ready);
                } else { pi = TS_VALID; pe = TS_VALID; }
                const unsigned pm = __ballot_sync(0xffffffffu, has_p);
                const int first_p = pm ? (__ffs((int)pm) - 1) : 32;            // nearest predecessor holding an inclusive prefix
                uint64_t ci = 0, ce = 0;
                if (lane < first_p) { ci = (ag & ~TS_VALID) >> 40; ce = ag & ((1ull << 40) - 1); }
                else if (lane == first_p) { ci = pi & ~TS_VALID; ce = pe & ~TS_VALID; }
#pragma unroll
                for (int o = 16; o; o >>= 1) { ci += __shfl_down_sync(0xffffffffu, ci, o); ce += __shfl_down_sync(0xffffffffu, ce, o); }
                excl_inf += __shfl_sync(0xffffffffu, ci, 0); excl_ev += __shfl_sync(0xffffffffu, ce, 0);
                if (pm) break;
                idx -= 32;
            }
        }
        if (lane == 0) {
            const uint64_t inc_inf = excl_inf + (agg >> 40), inc_ev = excl_ev + (agg & ((1ull << 40) - 1));
            *reinterpret_cast<volatile unsigned long long*>(st + 3 * tile + 1) = TS_VALID | inc_inf;
            *reinterpret_cast<volatile unsigned long long*>(st + 3 * tile + 2) = TS_VALID | inc_ev;
            s_excl = (excl_inf << 38) | excl_ev;
            if ((tile + 1) * TAIL_TILE >= S) class_counts[7] = (inc_inf << 38) | inc_ev;      // totals: S_inf, E
        }
    }
    __syncthreads();
    if (s < S) {
        const uint64_t ex = incl - mine + s_warp[warp];                 // exclusive within the tile
        const uint64_t i = (s_excl >> 38) + (ex >> 40);
        const uint64_t off = (s_excl & ((1ull << 38) - 1)) + (ex & ((1ull << 40) - 1));
        scan_out[s] = (i << 38) | off;
        if (mine) {
            inf_site[i] = (int32_t)s;
            AlleleMeta m1{}, m2{};
            m1.off = (int64_t)off; m1.n = (int32_t)n1; m1.flags = u1 ? PB_AF_IN_USE : 0;
            m2.off = (int64_t)(off + n1); m2.n = (int32_t)n2; m2.flags = u2 ? PB_AF_IN_USE : 0;
            alleles[2 * i] = m1;
            alleles[2 * i + 1] = m2;
            atomicMax(&s_nmax, max(n1, n2));
            atomicAdd(&s_ause, (unsigned int)((u1 ? 1 : 0) + (u2 ? 1 : 0)));
        }
    }
    __syncthreads();
    if (tid < 5 && s_cls[tid]) atomicAdd(class_counts + tid, (unsigned long long)s_cls[tid]);
    if (tid == 0 && s_ause) {
        atomicMax(class_counts + 5, (unsigned long long)s_nmax);
        atomicAdd(class_counts + 6, (unsigned long long)s_ause);
    }
}

// ---- CSR of extant events ------------------------------------------------------------------------------------------
template <bool XMODE>
__global__ void k1_scatter_events_kernel(const raw_rec_t* __restrict__ raw, const unsigned long long* __restrict__ begin,
                                         const unsigned long long* __restrict__ cursor, int64_t raw_cap,
                                         const int32_t* __restrict__ node_sample, const uint8_t* __restrict__ flags,
                                         const uint64_t* __restrict__ scan, const AlleleMeta* __restrict__ alleles,
                                         const uint64_t* __restrict__ xmask, uint32_t* __restrict__ csr_cursor, int32_t* __restrict__ csr,
                                         int64_t csr_cap, unsigned long long* __restrict__ cursor_copy) {
    const int64_t n = min((int64_t)*cursor, raw_cap), i0 = min((int64_t)*begin, n);
    if (blockIdx.x == 0 && threadIdx.x == 0) *cursor_copy = *cursor;   // the sweep's record count joins the tail's one readback
    for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const ev_rec_t r = load_rec<XMODE>(raw, i);
        if (!r.leaf || r.nomrca) continue;                // extant (leaf) events only (EXTANT_NODES_ONLY, HC.java:173)
        const int32_t smp = node_sample[r.node];
        if (smp < 0) continue;                            // leaf without a pheno row: never tallied (HC.java:3743)
        uint64_t m = r.mask;
        while (m) {
            const int b = __ffsll((long long)m) - 1;
            m &= m - 1;
            const uint32_t site = r.wcol * 64u + (uint32_t)b;
            const uint8_t fl = flags[site];
            if (!(fl & PB_SF_INF)) continue;
            const uint32_t base = event_base<XMODE>(r, b, xmask);
            const int a = (base == ((fl >> PB_SF_A1_CODE_SHIFT) & 3u)) ? 0 : 1;
            // allele in use, and its bucket not emptied by the <2 rule (in_use with min_hcount 0 keeps n = 0)
            if (!(fl & (a == 0 ? PB_SF_A1_USE : PB_SF_A2_USE)) || !(fl & (a == 0 ? PB_SF_A1_ALIVE : PB_SF_A2_ALIVE))) continue;
            const int64_t slot = 2 * (int64_t)(scan[site] >> 38) + a;
            const uint32_t k = atomicAdd(csr_cursor + slot, 1u);
            const int64_t o = alleles[slot].off + k;
            if (o < csr_cap) csr[o] = smp;               // (E beyond the buffer: pb_run_events grows it and repeats the pass)
        }
    }
}

// =================================================================================================
// The caller's padding bits of the last site word (S % 64 != 0) must never count as sites: they are cleared from the
// V plane once, when the tile holding word W-1 is uploaded, instead of being masked per row inside K1.
__global__ void k1_mask_last_word_kernel(uint64_t* __restrict__ V, int32_t N, int64_t Wp, int64_t w_last, uint64_t mask) {
    const int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N) V[(int64_t)v * Wp + w_last] &= mask;
}

int pbk_mask_last_word(pb_ctx* ctx) {
    if (!(ctx->S & 63) || ctx->N == 0) return PB_OK;
    const uint64_t mask = (1ull << (ctx->S & 63)) - 1;
    k1_mask_last_word_kernel<<<(ctx->N + 255) / 256, 256, 0, ctx->upl>>>(ctx->d_V, ctx->N, ctx->Wp, ctx->W - 1, mask);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// ---- compact (biallelic) upload: expand X in place --------------------------------------------------------
// pb_put_planes_biallelic copied the tile's X plane into the B0 rows: B0 = A0 ^ (X & D0), B1 = A1 ^ (X & D1), and, when
// the caller sent no V plane, V = all sites valid (padding bits of site word W-1 excluded).  One thread owns one word
// column (its four masks stay in registers) and walks a slice of the rows: 8 B read + 16/24 B written per node word.
__global__ void k1_expand_biallelic_kernel(uint64_t* __restrict__ B0, uint64_t* __restrict__ B1, uint64_t* __restrict__ V, int32_t N,
                                           int64_t Wp, int64_t w0, int64_t nw, const uint64_t* __restrict__ colmask, int fill_v,
                                           int64_t w_last, uint64_t last_mask) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nw) return;
    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(colmask + 4 * c), d = *reinterpret_cast<const ulonglong2*>(colmask + 4 * c + 2);
    const int64_t w = w0 + c;
    const uint64_t full = w == w_last ? last_mask : ~0ull;
    for (int32_t v = blockIdx.y; v < N; v += gridDim.y) {
        const int64_t o = (int64_t)v * Wp + w;
        const uint64_t x = B0[o];
        B0[o] = a.x ^ (x & d.x);
        B1[o] = a.y ^ (x & d.y);
        if (fill_v) V[o] = full;
    }
}

int pbk_expand_biallelic(pb_ctx* ctx, int64_t word_begin, int64_t n_words, const uint64_t* d_colmask, bool fill_v) {
    const uint64_t last_mask = (ctx->S & 63) ? (1ull << (ctx->S & 63)) - 1 : ~0ull;
    const unsigned gx = (unsigned)((n_words + 127) / 128);
    const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ctx->N, ((int64_t)ctx->sm_count * 16 + gx - 1) / gx));
    k1_expand_biallelic_kernel<<<dim3(gx, gy), 128, 0, ctx->upl>>>(ctx->d_B0, ctx->d_B1, ctx->d_V, ctx->N, ctx->Wp, word_begin, n_words,
                                                                     d_colmask, fill_v ? 1 : 0, ctx->W - 1, last_mask);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// X-mode helpers: padding bits of site word W-1 never count (X and the valid-site word), and the promotion to three planes
__global__ void k1x_mask_valid_kernel(uint64_t* __restrict__ xvalid, int64_t w_last, uint64_t mask) { xvalid[w_last] &= mask; }

int pbk_mask_last_word_x(pb_ctx* ctx) {
    if (!(ctx->S & 63) || ctx->N == 0) return PB_OK;
    const uint64_t mask = (1ull << (ctx->S & 63)) - 1;
    k1_mask_last_word_kernel<<<(ctx->N + 255) / 256, 256, 0, ctx->upl>>>(ctx->d_X, ctx->N, ctx->Wp, ctx->W - 1, mask);
    k1x_mask_valid_kernel<<<1, 1, 0, ctx->upl>>>(ctx->d_xvalid, ctx->W - 1, mask);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

__global__ void k1x_promote_kernel(const uint64_t* __restrict__ X, const uint64_t* __restrict__ xmask, const uint64_t* __restrict__ xvalid,
                                   uint64_t* __restrict__ B0, uint64_t* __restrict__ B1, uint64_t* __restrict__ V, int32_t N, int64_t Wp, int64_t W) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(xmask + 4 * w), d = *reinterpret_cast<const ulonglong2*>(xmask + 4 * w + 2);
    const uint64_t valid = xvalid[w];
    for (int32_t v = blockIdx.y; v < N; v += gridDim.y) {
        const int64_t o = (int64_t)v * Wp + w;
        const uint64_t x = X[o];
        B0[o] = a.x ^ (x & d.x);
        B1[o] = a.y ^ (x & d.y);
        V[o] = valid;
    }
}

int pbk_promote_planes(pb_ctx* ctx) {
    const unsigned gx = (unsigned)((ctx->W + 127) / 128);
    const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ctx->N, ((int64_t)ctx->sm_count * 16 + gx - 1) / gx));
    k1x_promote_kernel<<<dim3(gx, gy), 128, 0, ctx->upl>>>(ctx->d_X, ctx->d_xmask, ctx->d_xvalid, ctx->d_B0, ctx->d_B1, ctx->d_V, ctx->N,
                                                             ctx->Wp, ctx->W);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// one-plane sweep of the column blocks [block0, block0 + n_blocks).  first: a new pass starts (record cursor reset, start
// mark); last: the closing mark.  Tiles swept ahead of pb_run_events (eager_sweep in pb_api.cu) append to the same list.
int pbk_events_range(pb_ctx* ctx, int64_t block0, int64_t n_blocks, bool first, bool last) {
    KXParams q;
    q.col_block0 = (int32_t)block0;
    q.ops = ctx->d_ops_x; q.chunk_op = ctx->d_chunk_op_x;
    q.raw = reinterpret_cast<xrec_t*>(ctx->d_raw); q.raw_cursor = ctx->d_raw_cursor; q.raw_cap = (unsigned long long)ctx->raw_cap;
    const size_t smem = (size_t)KX_SMEM_TOTAL;
    if (!ctx->k1x_attr_set) {     // a function attribute belongs to the device: once per context, not once per process
        PB_CUDA(cudaFuncSetAttribute(k1x_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->k1x_attr_set = true;
    }
    if (first) PB_CUDA(cudaMemsetAsync(ctx->d_raw_cursor, 0, 8, ctx->stream));
    if (first || last) PB_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));     // (a closing launch times itself only)
    if (first) pb_trace_mark(ctx, "start");
    if (n_blocks > 0) {
        dim3 grid((unsigned)n_blocks, (unsigned)ctx->n_chunks_x);
        k1x_events_kernel<<<grid, PB_K1_THREADS, smem, ctx->stream>>>(ctx->tmap_x, q);
        ctx->tm.n_launches += 1;
    }
    if (last) { pb_trace_mark(ctx, "k1x"); PB_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream)); }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

static int pbk_events_x(pb_ctx* ctx) {
    return pbk_events_range(ctx, 0, (ctx->W + PB_KX_COLS - 1) / PB_KX_COLS, true, true);
}

int pbk_events(pb_ctx* ctx) {
    if (ctx->plane_mode == PB_PLANES_X) return pbk_events_x(ctx);
    K1Params p;
    p.W = ctx->W;
    p.ops = ctx->d_ops; p.chunk_op = ctx->d_chunk_op;
    p.census = ctx->d_census;
    p.raw = ctx->d_raw; p.raw_cursor = ctx->d_raw_cursor; p.raw_cap = (unsigned long long)ctx->raw_cap;
    size_t smem = (size_t)K1_SMEM_TOTAL;
    if (const char* e = getenv("PB_K1_SMEM_PAD_KB")) smem += (size_t)atoi(e) * 1024;   // experiment knob: lowers blocks/SM
    if (!ctx->k1_attr_set) {      // a function attribute belongs to the device: once per context, not once per process
        PB_CUDA(cudaFuncSetAttribute(k1_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->k1_attr_set = true;
    }
    dim3 grid((unsigned)((ctx->W + PB_K1_COLS - 1) / PB_K1_COLS), (unsigned)ctx->n_chunks);
    PB_CUDA(cudaMemsetAsync(ctx->d_raw_cursor, 0, 8, ctx->stream));
    PB_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    pb_trace_mark(ctx, "start");
    k1_events_kernel<<<grid, PB_K1_THREADS, smem, ctx->stream>>>(ctx->tmap, p);
    pb_trace_mark(ctx, "k1");
    PB_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    ctx->tm.n_launches += 1;
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// Everything after the K1 sweep, enqueued without a host round trip: the raw list length, S_inf and E stay on the
// device, the compaction outputs are sized for the worst case (every site informative, every raw event extant).
// One readback at the end: class counters, n_max, #alleles in use, scan totals, K1's cursor.
// Buffers of the tail (worst-case sizes) and the zero-fill of its counters.  The 56 MB of memsets at C2 do not depend on
// K1, so they go to the second stream BEFORE the sweep is enqueued and run underneath it; the tail waits on ev_zero_done.
int pbk_prepare_tail(pb_ctx* ctx) {
    const int64_t S = ctx->S;
    if (S > ctx->inf_cap) {
        cudaFree(ctx->d_inf_site); cudaFree(ctx->d_alleles); cudaFree(ctx->d_csr_cursor); cudaFree(ctx->d_pobs); cudaFree(ctx->d_r);
        ctx->d_inf_site = nullptr; ctx->d_alleles = nullptr; ctx->d_csr_cursor = nullptr; ctx->d_pobs = nullptr; ctx->d_r = nullptr;
        PB_CUDA(cudaMalloc(&ctx->d_inf_site, (size_t)S * 4));
        PB_CUDA(cudaMalloc(&ctx->d_alleles, (size_t)S * 2 * sizeof(AlleleMeta)));
        PB_CUDA(cudaMalloc(&ctx->d_csr_cursor, (size_t)S * 2 * 4));
        PB_CUDA(cudaMalloc(&ctx->d_pobs, (size_t)S * 2 * 8));
        PB_CUDA(cudaMalloc(&ctx->d_r, (size_t)S * 2 * 4));
        ctx->inf_cap = S;
    }
    if (std::max(ctx->raw_cap, ctx->csr_want) > ctx->csr_cap) {
        cudaFree(ctx->d_csr); ctx->d_csr = nullptr;
        ctx->csr_cap = std::max(ctx->raw_cap, ctx->csr_want);
        PB_CUDA(cudaMalloc(&ctx->d_csr, (size_t)ctx->csr_cap * 4));
    }
    {
        const int64_t need = 1 + 3 * ((S + TAIL_TILE - 1) / TAIL_TILE);     // ticket + per-tile {aggregate, prefix inf, prefix ev}
        if (need > ctx->scan_state_cap) {
            if (ctx->d_scan_state) cudaFree(ctx->d_scan_state);
            ctx->d_scan_state = nullptr; ctx->scan_state_cap = 0;
            PB_CUDA(cudaMalloc(&ctx->d_scan_state, (size_t)need * 8));
            ctx->scan_state_cap = need;
        }
        PB_CUDA(cudaMemsetAsync(ctx->d_scan_state, 0, (size_t)need * 8, ctx->stream2));
        PB_CUDA(cudaMemsetAsync(ctx->d_class_counts, 0, 12 * 8, ctx->stream2));   // words 12.. belong to pb_run_assoc
    }
    PB_CUDA(cudaMemsetAsync(ctx->d_cnt, 0, (size_t)S * PB_CNT_ALLOC_STRIDE * 4, ctx->stream2));   // counters + the one-plane census sum
    PB_CUDA(cudaMemsetAsync(ctx->d_csr_cursor, 0, (size_t)S * 2 * 4, ctx->stream2));
    PB_CUDA(cudaEventRecord(ctx->ev_zero_done, ctx->stream2));
    return PB_OK;
}

// The tail over the records [*rec_begin, *rec_end) and the TAIL_TILE-site tiles [tile0, tile0 + n_tiles) — the whole pass, or (one-plane
// mode) what one uploaded column tile's sweep appended: record counts -> per-site finalisation + compaction -> CSR of extant events.
// The single-pass scan hands out tile ids by ticket, so consecutive launches simply continue where the previous one stopped.
static int tail_launch(pb_ctx* ctx, const unsigned long long* rec_begin, const unsigned long long* rec_end, int64_t n_tiles, unsigned ev_blocks) {
    const int64_t S = ctx->S;
    const int T = 256;
    cudaStream_t st = ctx->stream;
    const bool xmode = ctx->plane_mode == PB_PLANES_X;
    const uint64_t* xmask = xmode ? ctx->d_xmask : nullptr;
    int32_t* xcensus = reinterpret_cast<int32_t*>(ctx->d_cnt + S * PB_CNT_STRIDE);
    int kt_used = 1;
    while ((1ll << kt_used) <= (int64_t)ctx->L) kt_used++;
    if (xmode) {
        // one-plane mode: the leaf census is a by-product of the record pass (no dense counters, nothing to reduce)
        k1_count_events_kernel<true><<<ev_blocks, T, 0, st>>>(ctx->d_raw, rec_begin, rec_end, ctx->raw_cap, ctx->d_node_sample, ctx->d_leaves_below,
                                                              xmask, ctx->d_cnt, xcensus, ctx->d_class_counts + 9);
        pb_trace_mark(ctx, "count_events");
        ctx->tm.n_launches += 1;
        if (n_tiles > 0) {
            k1_finalize_compact_kernel<true><<<(unsigned)n_tiles, TAIL_TILE, 0, st>>>(
                nullptr, xcensus, kt_used, ctx->W, S, ctx->d_X, ctx->d_xmask, ctx->d_xvalid, ctx->Wp, ctx->root, ctx->L, ctx->d_cnt,
                ctx->cfg.min_hcount, ctx->cfg.site_offset, ctx->d_sites, ctx->d_site_flags, ctx->d_site_scan, ctx->d_inf_site, ctx->d_alleles,
                ctx->d_scan_state, ctx->d_class_counts);
            ctx->tm.n_launches += 1;
        }
    } else {
        k1_count_events_kernel<false><<<ev_blocks, T, 0, st>>>(ctx->d_raw, rec_begin, rec_end, ctx->raw_cap, ctx->d_node_sample, nullptr, nullptr,
                                                               ctx->d_cnt, nullptr, ctx->d_class_counts + 9);
        pb_trace_mark(ctx, "count_events");
        k1_census_reduce_kernel<<<dim3((unsigned)((ctx->W + 127) / 128), 4), 128, 0, st>>>(ctx->d_census, ctx->n_flush, ctx->W, ctx->d_census_total);
        pb_trace_mark(ctx, "census_reduce");
        k1_finalize_compact_kernel<false><<<(unsigned)n_tiles, TAIL_TILE, 0, st>>>(
            ctx->d_census_total, nullptr, kt_used, ctx->W, S, ctx->d_B0, ctx->d_B1, ctx->d_V, ctx->Wp, ctx->root, ctx->L, ctx->d_cnt, ctx->cfg.min_hcount,
            ctx->cfg.site_offset, ctx->d_sites, ctx->d_site_flags, ctx->d_site_scan, ctx->d_inf_site, ctx->d_alleles, ctx->d_scan_state,
            ctx->d_class_counts);
        ctx->tm.n_launches += 3;
    }
    pb_trace_mark(ctx, "finalize+compact");
    PB_CUDA(cudaGetLastError());
    uint64_t* scan_out = ctx->d_site_scan;
    if (xmode)
        k1_scatter_events_kernel<true><<<ev_blocks, T, 0, st>>>(ctx->d_raw, rec_begin, rec_end, ctx->raw_cap, ctx->d_node_sample, ctx->d_site_flags,
                                                               scan_out, ctx->d_alleles, xmask, ctx->d_csr_cursor, ctx->d_csr, ctx->csr_cap, ctx->d_class_counts + 8);
    else
        k1_scatter_events_kernel<false><<<ev_blocks, T, 0, st>>>(ctx->d_raw, rec_begin, rec_end, ctx->raw_cap, ctx->d_node_sample, ctx->d_site_flags,
                                                                scan_out, ctx->d_alleles, xmask, ctx->d_csr_cursor, ctx->d_csr, ctx->csr_cap, ctx->d_class_counts + 8);
    pb_trace_mark(ctx, "scatter_events");
    ctx->tm.n_launches += 1;
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// One-plane mode, streamed upload: the tail of the column tile [word_begin, word_begin + n_words) that eager_sweep (pb_api.cu) has just
// swept, enqueued behind that sweep — it runs underneath the next tile's copy like the sweep does — and, when the host has said
// where the per-site records go (pb_expect_site_output), their copy back on the download stream: PCIe is full duplex and the
// device-to-host direction is idle during the upload.  pb_run_events then finds only the last tile's tail left.
// Tiles arrive in column order and are cut at multiples of 256 site words, so their sites are whole TAIL_TILE tiles.
int pbk_tail_tile(pb_ctx* ctx, int64_t word_begin, int64_t n_words) {
    if (ctx->tail_tiles >= PB_MAX_TAIL_TILES || word_begin != ctx->tailed_words) return PB_OK;   // (the rest is left to pb_run_events)
    cudaStream_t st = ctx->stream;
    const int t = ctx->tail_tiles;
    if (!ctx->d_rec_range) {
        PB_CUDA(cudaMalloc(&ctx->d_rec_range, (PB_MAX_TAIL_TILES + 1) * sizeof(unsigned long long)));
        PB_CUDA(cudaMemsetAsync(ctx->d_rec_range, 0, (PB_MAX_TAIL_TILES + 1) * sizeof(unsigned long long), st));    // [0] = 0 for good (stream-ordered: the streams do not wait for the legacy stream)
    }
    // the record cursor as this tile's sweep left it = the end of its record range
    PB_CUDA(cudaMemcpyAsync(ctx->d_rec_range + t + 1, ctx->d_raw_cursor, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    if (t == 0) PB_CUDA(cudaStreamWaitEvent(st, ctx->ev_zero_done, 0));     // pbk_prepare_tail's zero fills (second stream)
    const int64_t s0 = word_begin * 64, s1 = std::min<int64_t>(ctx->S, (word_begin + n_words) * 64);
    const int64_t n_tiles = (s1 - s0 + TAIL_TILE - 1) / TAIL_TILE;
    int rc = tail_launch(ctx, ctx->d_rec_range + t, ctx->d_rec_range + t + 1, n_tiles, (unsigned)ctx->sm_count * 4);
    if (rc) return rc;
    ctx->tail_tiles = t + 1;
    ctx->tailed_words = word_begin + n_words;
    if (ctx->out_sites && ctx->sent_sites == s0) {
        PB_CUDA(cudaEventRecord(ctx->ev_tail_tile, st));
        PB_CUDA(cudaStreamWaitEvent(ctx->dl, ctx->ev_tail_tile, 0));
        PB_CUDA(cudaMemcpyAsync(ctx->out_sites + s0, ctx->d_sites + s0, (size_t)(s1 - s0) * sizeof(pb_site_out), cudaMemcpyDeviceToHost, ctx->dl));
        ctx->sent_sites = s1;
    }
    return PB_OK;
}

// tailed_words: site words whose tail ran ahead of this call (pbk_tail_tile), 0 = none
int pbk_finalize_events(pb_ctx* ctx, pb_class_counts* counts, unsigned long long* cursor_out, int64_t tailed_words) {
    const int64_t S = ctx->S;
    const int T = 256;
    cudaStream_t st = ctx->stream;
    const unsigned ev_blocks = (unsigned)std::min<int64_t>((ctx->raw_cap + T - 1) / T, (int64_t)ctx->sm_count * 16);
    PB_CUDA(cudaStreamWaitEvent(st, ctx->ev_zero_done, 0));   // counters, cursors, scan state, class counts zeroed underneath K1 (pbk_prepare_tail)
    pb_trace_mark(ctx, "headstart+memsets");
    if (!ctx->d_rec_range) {
        PB_CUDA(cudaMalloc(&ctx->d_rec_range, (PB_MAX_TAIL_TILES + 1) * sizeof(unsigned long long)));
        PB_CUDA(cudaMemsetAsync(ctx->d_rec_range, 0, (PB_MAX_TAIL_TILES + 1) * sizeof(unsigned long long), st));
    }
    const int64_t all_tiles = (S + TAIL_TILE - 1) / TAIL_TILE;
    const int64_t done_tiles = tailed_words > 0 ? std::min<int64_t>(all_tiles, (tailed_words * 64 + TAIL_TILE - 1) / TAIL_TILE) : 0;
    const unsigned long long* rec_begin = ctx->d_rec_range + (tailed_words > 0 ? ctx->tail_tiles : 0);
    {
        int rc = tail_launch(ctx, rec_begin, ctx->d_raw_cursor, all_tiles - done_tiles, tailed_words > 0 ? (unsigned)ctx->sm_count * 4 : ev_blocks);
        if (rc) return rc;
    }
    unsigned long long h[16];
    PB_CUDA(cudaMemcpyAsync(h, ctx->d_class_counts, sizeof(h), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    PB_CUDA(cudaGetLastError());
    pb_trace_dump(ctx, "events");
    *cursor_out = h[8];
    ctx->n_raw = (int64_t)h[9];                       // events = set bits over all records (exact only when nothing overflowed)
    ctx->S_inf = (int64_t)(h[7] >> 38);
    ctx->E = (int64_t)(h[7] & ((1ull << 38) - 1));
    ctx->n_max = (int32_t)h[5];
    ctx->A_inuse = (int64_t)h[6];
    if (counts) {
        counts->n_no_allele = (int64_t)h[0]; counts->n_mono = (int64_t)h[1]; counts->n_bi = (int64_t)h[2];
        counts->n_tri = (int64_t)h[3]; counts->n_quad = (int64_t)h[4]; counts->n_raw_events = (int64_t)h[9];
    }
    return PB_OK;
}
