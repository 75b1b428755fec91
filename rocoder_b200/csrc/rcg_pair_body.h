// rcg_pair_body.h -- W = 65536: the body of the CTA-pair stretch kernel (rcg_frame_pair.h) and its barrier
// flavours.  Shared by rcg_kernels.cu and by the translation unit NVRTC builds when a bin-local frequency kernel
// is fused in (rcg_jit_entry.h), so the cluster operations are spelled in PTX instead of cooperative_groups
// (NVRTC sees no toolkit headers).  Device code only.
#pragma once
#include "rcg_frame_pair.h"
#include "rcg_kernel_body.h"

namespace rcg {

__device__ __forceinline__ unsigned cluster_cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// generic address of `p` (a shared-memory address of this CTA) in the shared memory of CTA `rank` of the cluster
template <class T>
__device__ __forceinline__ T* cluster_map_shared(T* p, unsigned rank) {
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)p), "r"(rank));
    return reinterpret_cast<T*>(out);
}
// all threads of all CTAs of the cluster; release / acquire: shared-memory writes before it are visible to
// distributed-shared-memory reads after it
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 32-bit shared-window address of a shared-memory object of this CTA / of the same object in CTA `rank`
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_addr(unsigned cta_addr, unsigned rank) {
    unsigned out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(cta_addr), "r"(rank));
    return out;
}

struct PairDeviceExec {
    int r, t;
    unsigned mbar = 0;         // this CTA's mbarrier: counts the bytes the peer's bulk copy delivers into my zone
    unsigned peer_mbar = 0;    // the peer's, as a shared::cluster address
    unsigned peer_zone = 0;    // the peer's landing zone, shared::cluster address
    unsigned peer_buf = 0;     // the peer's transform buffer, shared::cluster address
    unsigned parity = 0;
    // Both CTAs hand each other kPairLandCf values: one thread arms its own mbarrier with the byte count it expects
    // and starts ONE bulk asynchronous copy (TMA engine: no thread moves data) from its transform buffer into the
    // peer's landing zone; the copy signals the peer's mbarrier as the bytes arrive.  Everybody then waits for the
    // bytes of its own zone only.
    __device__ __forceinline__ void exchange(cf* const (&bufs)[2], cf* const (&)[2], int n_cf) {
        const unsigned bytes = (unsigned)n_cf * (unsigned)sizeof(cf);
        // the passes wrote the buffer through the generic proxy; the copy reads it through the async proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            const unsigned src = smem_addr(pair_push_source(pair_buf(bufs, r), r));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(peer_zone), "r"(src), "r"(bytes), "r"(peer_mbar) : "memory");
        }
        asm volatile(
            "{\n\t.reg .pred p;\n"
            "RCG_MBAR_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra RCG_MBAR_DONE;\n\t"
            "bra RCG_MBAR_WAIT;\n"
            "RCG_MBAR_DONE:\n\t}" ::"r"(mbar), "r"(parity) : "memory");
        parity ^= 1u;
    }
    template <class F>
    __device__ __forceinline__ void par(F f) {
        f(r, t);
        __syncthreads();
    }
    template <class F>
    __device__ __forceinline__ void par_steps(int n, F f) {
        for (int s = 0; s < n; s++) {
            f(s, r, t);
            __syncwarp();
        }
        __syncthreads();
    }
    // a phase whose threads also store into the peer's buffer (store_peer): the cluster barrier (release / acquire at
    // cluster scope) orders those stores before anything either CTA does next
    template <class F>
    __device__ __forceinline__ void par_pair(F f) {
        f(r, t);
        cluster_barrier();
    }
    // v into slot `slot` of the peer CTA's transform buffer (posted store through distributed shared memory)
    __device__ __forceinline__ void store_peer(cf* const (&)[2], int, int slot, cf v) {
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(peer_buf + (unsigned)slot * (unsigned)sizeof(cf)),
                     "f"(v.x), "f"(v.y) : "memory");
    }
    template <class F>
    __device__ __forceinline__ void par_cluster(F f) {
        f(r, t);
        __threadfence();     // carried half windows travel through global memory
        cluster_barrier();
    }
    __device__ __forceinline__ void cluster_sync() { cluster_barrier(); }
    // the two halves of one barrier, issued apart: whatever lies between them overlaps the peer's progress
    __device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
    __device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
};

// Shared memory of a CTA of the pair: transform buffer, landing zone, and the small twiddle tables -- those of the
// local passes >= 1 (1264 values) and the two-level table of the cross stage (320): the CTA holds one frame with 16
// warps, so every table load that goes to L2 is a bubble nothing hides (ncu: 10 % of the stall samples before).
// Pass 0's table (8192 values, read once per element, unit stride) stays in global memory.
constexpr int kPairTwSmemFirst = tw_offset(kPairLog2HL, 1);
constexpr int kPairTwSmemCf = tw_offset(kPairLog2HL, make_plan(kPairLog2HL).npasses) - kPairTwSmemFirst;
constexpr size_t kPairSmemBytes =
    (size_t)(FrameCfg<kPairLog2HL>::SMEM_CF + kPairLandCf + kPairTwSmemCf + kPairTwLen) * sizeof(cf) + 16    // + mbarrier
    + (FrameCfg<kPairLog2HL>::TW_TMEM ? 16 : 0);                                                              // + Tensor Memory slot

__device__ __forceinline__ void stretch_pair_body(const StretchParams& prm_in, const PairTables& pt_in, int n_items,
                                                  unsigned char* smem_raw) {
    const int r = (int)cluster_cta_rank();
    cf* S = reinterpret_cast<cf*>(smem_raw);
    cf* zone = S + FrameCfg<kPairLog2HL>::SMEM_CF;             // landing zone behind the transform buffer
    cf* tw_s = zone + kPairLandCf;
    cf* cross_s = tw_s + kPairTwSmemCf;
    if constexpr (!FrameCfg<kPairLog2HL>::TW_TMEM)
        for (int i = threadIdx.x; i < kPairTwSmemCf; i += blockDim.x) tw_s[i] = __ldg(prm_in.tw + kPairTwSmemFirst + i);
    for (int i = threadIdx.x; i < kPairTwLen; i += blockDim.x) cross_s[i] = __ldg(pt_in.cross_tw + i);
    __syncthreads();
    // the pass twiddles of the local transform: each thread's Tensor Memory row (rcg_common.h), or the tables above
    uint32_t* tm_slot = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(cross_s + kPairTwLen) + 16);
    StretchParams prm = with_tmem_twiddles<kPairLog2HL>(prm_in, tm_slot);
    prm.tw0 = prm_in.tw;
    prm.tw = tw_s - kPairTwSmemFirst;     // only passes >= 1 index it (pair_fft_range hands pass 0 prm.tw0)
    PairTables pt = pt_in;
    pt.cross_tw = cross_s;
    cf* bufs[2] = {S, S};                                      // only bufs[r] is touched: transforms are local
    cf* land[2] = {zone, zone};                                // only land[r] is read; the peer's zone is written by the copy
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(cross_s + kPairTwLen);
    PairDeviceExec ex{r, (int)threadIdx.x};
    ex.mbar = smem_addr(mbar);
    ex.peer_mbar = cluster_addr(ex.mbar, (unsigned)(r ^ 1));
    ex.peer_zone = cluster_addr(smem_addr(zone), (unsigned)(r ^ 1));
    ex.peer_buf = cluster_addr(smem_addr(S), (unsigned)(r ^ 1));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ex.mbar) : "memory");   // one arrival per phase: my own expect_tx
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_barrier();   // both mbarriers exist before either CTA can be sent bytes
    const int cid = blockIdx.x / 2, ncl = gridDim.x / 2;
    float* scratch = pt.scratch + (long long)cid * 3 * kPairH;
    for (int item = cid; item < n_items; item += ncl) run_work_item_pair(ex, bufs, land, prm, pt, item, scratch, true);
    cluster_barrier();  // no CTA may exit while its peer can still write into its shared memory
    release_tmem_twiddles<kPairLog2HL>(tm_slot);
}

}  // namespace rcg
