// One-shot all-reduce(sum) of a few hundred doubles over NVLink peer memory, fused into the kernel that produces them.
//
// The ELBO step (main_custom_training.py:199-214, 252-256) shards its Monte-Carlo samples over one process per GPU;
// what the ranks exchange per step is 3 + 4B doubles (2 KB at B = 64): latency, not bandwidth.  Instead of a separate
// collective after the reduction kernel, the reduction kernel itself STORES its partial sums into a mailbox in every
// peer's memory (plain P2P stores through NVLink / NVSwitch), raises a sequence flag there, waits for the peers' flags
// in its own mailbox and adds the world's partials in rank order -- every rank computes bit-identical totals,
// independent of arrival order.
//
// Mailbox of one rank (device memory of that rank, mapped into every peer through CUDA IPC or, for handles of one
// process, used by address):   double  data[2][world][cap]     parity-double-buffered slots, one per source rank
//                              uint64  flag[2][world]          sequence number of the call the slot belongs to
// Call number q (1, 2, ...) uses parity q & 1.  A rank can start call q + 2 (same parity) only after call q + 1, which
// needs every peer's flag of call q + 1, which a peer raises only after it has finished reading call q: two parities
// are enough, no slot is overwritten while it is being read.
#pragma once
#include <cstdint>

constexpr int kPeerMaxWorld = 16;

struct PeerCtx {
    int rank = 0, world = 0, cap = 0;
    double *mail[kPeerMaxWorld] = {};     // mailbox base of every rank as mapped HERE (mail[rank] = this rank's own)
    unsigned long long *seq = nullptr;    // local: number of completed calls
    unsigned int *arrived = nullptr;      // local: blocks of the current kernel that have pushed their part
    int *err = nullptr;                   // local: 1 after a wait timed out (a peer never arrived)
    unsigned long long timeout_ns = 0;
};

__device__ __forceinline__ unsigned long long *peer_flags(const PeerCtx &P, int r) {
    return (unsigned long long *)(P.mail[r] + (size_t)2 * P.world * P.cap);
}
__device__ __forceinline__ unsigned long long peer_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// slot of this rank, parity of the call in flight, in rank r's mailbox
__device__ __forceinline__ double *peer_slot(const PeerCtx &P, unsigned long long q, int r) {
    return P.mail[r] + ((size_t)(q & 1) * P.world + P.rank) * P.cap;
}

// Every block that contributes calls peer_push for its own entries (value v -> index i of this rank's slot in every
// mailbox), then ALL its threads call peer_finish.  The last block to arrive (of nblocks) signals the peers, waits
// for them and writes the rank-ordered totals of entries [0, n) to out.  Returns true in the block that finished.
__device__ __forceinline__ void peer_push(const PeerCtx &P, unsigned long long q, int i, double v) {
    for (int r = 0; r < P.world; ++r) peer_slot(P, q, r)[i] = v;
}

__device__ __forceinline__ bool peer_finish(const PeerCtx &P, unsigned long long q, int nblocks, int n, double *out) {
    __shared__ unsigned int s_last;
    const int tid = threadIdx.x;
    __threadfence_system();  // this thread's pushes are visible system-wide before the arrival below
    __syncthreads();
    if (tid == 0) s_last = (nblocks == 1) ? 1u : (atomicAdd(P.arrived, 1u) == (unsigned)nblocks - 1u);
    __syncthreads();
    if (!s_last) return false;
    __threadfence_system();  // cumulativity: the other blocks' pushes (observed through the counter) before the flags
    const int slot = (int)(q & 1) * P.world;
    if (tid < P.world) {
        unsigned long long *f = peer_flags(P, tid) + slot + P.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(q) : "memory");
        const unsigned long long *mine = peer_flags(P, P.rank) + slot + tid;
        const unsigned long long t0 = peer_globaltimer();
        unsigned long long seen;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
            if (seen >= q) break;
            if (peer_globaltimer() - t0 > P.timeout_ns) {
                *P.err = 1;
                break;
            }
        }
    }
    __syncthreads();
    const double *own = P.mail[P.rank] + (size_t)slot * P.cap;
    for (int i = tid; i < n; i += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < P.world; ++r) acc += __ldcv(own + (size_t)r * P.cap + i);  // rank order, never from L1
        out[i] = acc;
    }
    if (tid == 0) {
        *P.seq = q;
        *P.arrived = 0u;
    }
    return true;
}

// Stand-alone form: in-place all-reduce(sum) of buf[0, n) (one block).
__global__ void peer_allreduce_kernel(PeerCtx P, double *buf, int n) {
    const unsigned long long q = *P.seq + 1;
    __syncthreads();  // everyone has read the call number before thread 0 advances it
    for (int i = threadIdx.x; i < n; i += blockDim.x) peer_push(P, q, i, buf[i]);
    peer_finish(P, q, 1, n, buf);
}
