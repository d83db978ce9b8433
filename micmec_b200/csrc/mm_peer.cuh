// mm_peer.cuh - layout of the peer-visible control block of a z-slab (mm_comm.cu) and the system-scope accessors, shared
// with the tail of the marching kernel (mm_march2.cuh), which runs the mailbox all-reduce itself.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mm {

struct PeerCtl {
    unsigned long long halo_epoch, red_epoch;
    unsigned int halo_done, pad;
    unsigned long long fused_epoch;  // fused halo exchanges completed (k_halo_xy_fused)
    unsigned int fused_done, pad2;
};

constexpr size_t kPeerFlagBytes = 1024;
static inline size_t peer_mail_bytes(int P) { return sizeof(double) * 2 * P * 16; }
static inline size_t peer_inbox_off(int P) { return kPeerFlagBytes + ((peer_mail_bytes(P) + 1023) & ~(size_t)1023); }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace mm
