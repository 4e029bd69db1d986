// mirage_b200/csrc/common.cuh
//
// sm_100a building blocks shared by every kernel in libmirage_b200.so:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st),
// UMMA shared-memory + instruction descriptors, and small math helpers.
//
// Everything here is inline PTX; there is no CUTLASS/CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mb200 {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define MB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::mb200::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                              cudaGetErrorString(_e));                                        \
      return -2;                                                                              \
    }                                                                                         \
  } while (0)

#define MB_REQUIRE(cond, ...)                                                                 \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ::mb200::set_last_error(__VA_ARGS__);                                                   \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
  // make barrier inits visible to the async proxy (TMA / tcgen05.commit)
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// non-blocking probe (no hardware suspend): used by the attention MMA scheduler to poll several barriers
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: a protocol bug must never hang the GPU (a hung box is a lost box).  After ~4 s of
// spinning the kernel traps, which surfaces as a CUDA error on the host.  Build with
// -DMB_DEBUG_BARRIERS to also print which barrier timed out (the printf call costs registers and
// instruction-cache footprint at every wait site, so it is off by default).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {
#ifdef MB_DEBUG_BARRIERS
      printf("mirage_b200: mbarrier timeout block=(%d,%d,%d) thread=%d bar=0x%x parity=%u\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, smem_u32(bar), parity);
#endif
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// clusters / CTA pairs (cta_group::2)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared::cluster address of the same smem location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}

// TMA load whose completion bytes are credited to the mbarrier of the LEADER CTA of a CTA pair
// (bar_cluster_addr is a shared::cluster address); destination is this CTA's own smem.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m,
                                                 uint32_t bar_cluster_addr, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0),
      "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d_pair(void* smem_dst, const CUtensorMap* m,
                                                 uint32_t bar_cluster_addr, int32_t c0, int32_t c1,
                                                 int32_t c2, int32_t c3, int32_t c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0),
      "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}

__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// D[tmem of both CTAs] (+)= A * B with M = 256 split over the CTA pair; issued by ONE thread of the
// leader CTA.  A: 128 rows per CTA; B: N/2 rows per CTA (each tensor core reads both halves).
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_tf32_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// commit of the pair's MMAs: arrives on the mbarrier at this smem offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// proxies / fences
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  // generic-proxy smem writes -> visible to the async proxy (TMA store, UMMA operand reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2, int32_t c3,
                                            int32_t c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0,
                                             int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, ld/st
// ---------------------------------------------------------------------------------------------
// One full warp calls alloc; the TMEM base address lands in *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}

__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.  kind::f16 covers bf16/fp16 inputs with
// fp32 accumulate; kind::tf32 takes fp32 storage and rounds to tf32 inside the tensor core.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive (count 1) on an mbarrier once every previously issued tcgen05.mma of this thread retired.
// Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x (32 consecutive 32-bit columns): thread i of the warp reads TMEM lane (base_lane + i).
// The warp may only touch the lane quarter 32*(warp_id % 4).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// pointer form (v must index registers at compile time after unrolling)
__device__ __forceinline__ void tmem_ld_32x32b_x32_p(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
      "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
      "r"(v[15])
      : "memory");
}

// pointer form (v must index registers at compile time after unrolling)
__device__ __forceinline__ void tmem_st_32x32b_x16_p(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
      "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
      "r"(v[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
      "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
      "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]),
      "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
      "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (sm_100 "version 1" shared-memory matrix descriptor, 128-byte swizzle)
// ---------------------------------------------------------------------------------------------
// All operand tiles in this library are staged by TMA with CU_TENSOR_MAP_SWIZZLE_128B, i.e. as
// rows of 128 bytes whose 16-byte chunks are XOR-ed with (row % 8); 8 rows form a 1024-byte atom.
//
//  K-major operand  (reduction dim contiguous in global memory):
//      one smem row per M/N index, 64 bf16 (or 32 tf32) of K per row.
//      SBO = 1024 B (next group of 8 M/N rows); LBO unused.  Advancing K by one UMMA_K (32 B) inside
//      the 128-byte row = start address + 32 B.
//  MN-major operand (M/N dim contiguous in global memory):
//      one smem row per K index, 64 bf16 of M/N per row ("slab"); a tile wider than 64 along M/N
//      is several slabs, each slab_bytes apart.
//      LBO = slab_bytes (next 64 M/N), SBO = 1024 B (next group of 8 K rows).
//      Advancing K by one UMMA_K (16 rows) = start address + 2048 B.
//
// The same holds for 64-byte swizzle (rows of 64 bytes, 8-row atoms of 512 bytes) with every
// "128" above halved; it is used for head_dim = 32 attention operands.
constexpr uint64_t kDescVersion1 = 1ull << 46;
constexpr uint64_t kDescSwizzle128B = 2ull << 61;
constexpr uint64_t kDescSwizzle64B = 4ull << 61;

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes,
                                                   uint64_t swizzle = kDescSwizzle128B) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= kDescVersion1;
  d |= swizzle;
  return d;
}

enum : uint32_t { kFmtF16 = 0, kFmtBF16 = 1, kFmtTF32 = 2 };

// Instruction descriptor for kind::f16 / kind::tf32, fp32 accumulate.
// a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, uint32_t fmt,
                                                  uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4)            // c_format = f32
         | (fmt << 7)         // a_format
         | (fmt << 10)        // b_format
         | (a_mn << 15)       // a_major
         | (b_mn << 16)       // b_major
         | ((N >> 3) << 17)   // n_dim
         | ((M >> 4) << 24);  // m_dim
}

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// erf(x) by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32-exact for a bf16 result):
// one MUFU.RCP + one MUFU.EX2 + ~10 FMA-pipe instructions instead of erff()'s ~40.
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  // exp(-x^2) = 2^(-x^2 * log2 e)
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-ax * ax * 1.4426950408889634f));
  const float r = fmaf(-poly, e, 1.0f);
  return copysignf(r, x);
}

// exact (erf) GELU, as nn.GELU() in the reference (mirage/utils.py:143,156)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f));
}

// Forward GELU for the GEMM epilogue, one MUFU + 7 FMA-pipe slots per element (the A&S form above
// needs two MUFUs and ~16 slots, which made the fc1 epilogue, not the tensor pipe, the bound of
// that GEMM: ncu r01, 1442 us vs 790 us for the same tile loop with a plain bf16 epilogue).
//   erf-GELU(x) = 0.5 x (1 + erf(x / sqrt 2))  ~=  0.5 x (1 + tanh(x (a + b x^2 + c x^4)))
// a, b, c: least-squares fit to the exact erf form on [-8, 8], max |error| 3.0e-5 (oracle/fit_gelu.py);
// x^2 is clamped at 64 so the quartic never turns over (tanh is saturated there: 13.6).
// tanh.approx.f32 adds <= 2^-11 relative error in t, i.e. <= 2.5e-4 |x| absolute -- an order of
// magnitude below the bf16 rounding of the stored activation.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 64.0f);
  float p = fmaf(-3.58732362e-04f, x2, 3.70503451e-02f);
  p = fmaf(p, x2, 7.97458471e-01f);
  const float t = tanh_approx(x * p);
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

// two elements at once on the packed fp32x2 FMA pipe (Blackwell FMUL2 / FFMA2)
__device__ __forceinline__ void gelu_fast2(float& a, float& b) {
  float ta, tb;
  asm("{\n"
      ".reg .b64 x, x2, p, c2, c1, c0, hh, h, t;\n"
      ".reg .f32 lo, hi;\n"
      "mov.b64 x, {%2, %3};\n"
      "mul.rn.f32x2 x2, x, x;\n"
      "mov.b64 {lo, hi}, x2;\n"
      "min.f32 lo, lo, 0f42800000;\n"
      "min.f32 hi, hi, 0f42800000;\n"
      "mov.b64 x2, {lo, hi};\n"
      "mov.b64 c2, {0fB9BC143E, 0fB9BC143E};\n"   // -3.58732362e-04
      "mov.b64 c1, {0f3D17C21A, 0f3D17C21A};\n"   //  3.70503451e-02
      "mov.b64 c0, {0f3F4C263D, 0f3F4C263D};\n"   //  7.97458471e-01
      "mov.b64 hh, {0f3F000000, 0f3F000000};\n"   //  0.5
      "fma.rn.f32x2 p, c2, x2, c1;\n"
      "fma.rn.f32x2 p, p, x2, c0;\n"
      "mul.rn.f32x2 p, p, x;\n"
      "mov.b64 {lo, hi}, p;\n"
      "tanh.approx.f32 lo, lo;\n"
      "tanh.approx.f32 hi, hi;\n"
      "mov.b64 t, {lo, hi};\n"
      "mul.rn.f32x2 h, x, hh;\n"
      "fma.rn.f32x2 t, h, t, h;\n"
      "mov.b64 {%0, %1}, t;\n"
      "}\n"
      : "=f"(ta), "=f"(tb)
      : "f"(a), "f"(b));
  a = ta;
  b = tb;
}

// d/dx of the same tanh-form fit (consistent with gelu_fast2; max |error| vs the exact erf-GELU
// derivative 1.2e-4): g'(x) = 0.5 (1 + t) + 0.5 x (1 - t^2) u'(x),  t = tanh(u),  u = x (a + b x^2 + c x^4),
// u' = a + 3 b x^2 + 5 c x^4.  Where x^2 is clamped t = +-1 and the second term vanishes.
// One MUFU + ~7 FMA-pipe slots per element instead of erf + exp (3 MUFUs, ~25 slots).
__device__ __forceinline__ void gelu_fast_grad2(float& a, float& b) {
  float ga, gb;
  asm("{\n"
      ".reg .b64 x, x2, p, q, t, w, h, k;\n"
      ".reg .f32 lo, hi;\n"
      "mov.b64 x, {%2, %3};\n"
      "mul.rn.f32x2 x2, x, x;\n"
      "mov.b64 {lo, hi}, x2;\n"
      "min.f32 lo, lo, 0f42800000;\n"
      "min.f32 hi, hi, 0f42800000;\n"
      "mov.b64 x2, {lo, hi};\n"
      "mov.b64 k, {0fB9BC143E, 0fB9BC143E};\n"   // c
      "mov.b64 p, {0f3D17C21A, 0f3D17C21A};\n"   // b
      "fma.rn.f32x2 p, k, x2, p;\n"
      "mov.b64 k, {0f3F4C263D, 0f3F4C263D};\n"   // a
      "fma.rn.f32x2 p, p, x2, k;\n"                    // a + b x^2 + c x^4
      "mul.rn.f32x2 p, p, x;\n"                        // u
      "mov.b64 {lo, hi}, p;\n"
      "tanh.approx.f32 lo, lo;\n"
      "tanh.approx.f32 hi, hi;\n"
      "mov.b64 t, {lo, hi};\n"
      "mov.b64 k, {0fBAEB194E, 0fBAEB194E};\n"   // 5c
      "mov.b64 q, {0f3DE3A327, 0f3DE3A327};\n"   // 3b
      "fma.rn.f32x2 q, k, x2, q;\n"
      "mov.b64 k, {0f3F4C263D, 0f3F4C263D};\n"
      "fma.rn.f32x2 q, q, x2, k;\n"                    // u'
      "mov.b64 k, {0fBF800000, 0fBF800000};\n"
      "mul.rn.f32x2 w, t, k;\n"                        // -t
      "mov.b64 k, {0f3F800000, 0f3F800000};\n"
      "fma.rn.f32x2 w, w, t, k;\n"                     // 1 - t^2
      "mov.b64 k, {0f3F000000, 0f3F000000};\n"
      "mul.rn.f32x2 h, x, k;\n"                        // 0.5 x
      "mul.rn.f32x2 h, h, w;\n"                        // 0.5 x (1 - t^2)
      "fma.rn.f32x2 t, t, k, k;\n"                     // 0.5 t + 0.5
      "fma.rn.f32x2 t, h, q, t;\n"
      "mov.b64 {%0, %1}, t;\n"
      "}\n"
      : "=f"(ga), "=f"(gb)
      : "f"(a), "f"(b));
  a = ga;
  b = gb;
}

__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erf_as(x * 0.70710678118654752f));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.5f * x * x * 1.4426950408889634f));
  return fmaf(x * 0.3989422804014327f, e, cdf);
}

// fire-and-forget vector reduction into global memory (sm_90+): one L2 atomic transaction per 16 bytes
// instead of four scalar atomicAdds -- the split-K / direct-to-gradient-bucket wgrad epilogue
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// explicit shared-window accesses (keeps the compiler from emitting generic LD/ST for smem)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}

// 2^x on the MUFU unit; ex2.approx(-inf) = +0
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2): halves the FMA-pipe issue slots of the softmax
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1,
                                      float c0, float c1) {
  asm("{\n"
      ".reg .b64 ra, rb, rc, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "mov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n"
      ".reg .b64 ra, rb, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "add.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// host: TMA descriptor construction (cuTensorMapEncodeTiled through the runtime's driver entry
// point, so the library does not link libcuda directly)
// ---------------------------------------------------------------------------------------------
enum TmaDtype { kTmaBF16 = 0, kTmaF32 = 1 };

// dims / strides are innermost-first; strides_bytes has rank-1 entries (dim 1..rank-1).
// swizzle_bytes: 128 (default) or 64.
int make_tensor_map(CUtensorMap* out, const void* base, TmaDtype dtype, int rank,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    int swizzle_bytes = 128);

// SMs persistent kernels may occupy on the current device: the physical count minus what
// mb_set_sm_reserve() keeps free for concurrently running communication kernels (NCCL).
int sm_count();

// "Serpentine" row order between consecutive kernels.  Activations are 50 MB - 1 GB, the L2 holds 126 MB: a
// consumer that walks its rows in the SAME direction as its producer finds nothing of what the producer wrote
// (cyclic access over a buffer larger than an LRU cache), one that walks in the OPPOSITE direction starts with the
// producer's last ~100 MB still on chip.  The library keeps the direction in which the most recent participating
// kernel walked its rows; take_direction() flips it and returns the direction the kernel being launched must walk
// (+1 upwards, -1 downwards); kernels with a fixed order call note_direction(+1).  Host-side state, one stream
// assumed (as everywhere in this library).
// Measured (profiles/r02_ab_experiments.txt): the full alternation is NEUTRAL against everything-upwards at cfg 2
// and cfg 4, while reversing ONLY the LayerNorm kernels (GEMMs and attention upwards) gives +0.3-1.3 % -- so that
// is the default, and MB_SERPENTINE=1 switches the full alternation on.
bool serpentine_enabled();
int take_direction();
void note_direction(int dir);

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A kernel launched through launch_k() may start while its stream
// predecessor is still draining: its CTAs become resident as SM resources free up and run their private
// prologue (barrier init, TMEM allocation, descriptor prefetch, index math).  pdl_wait() blocks until the
// predecessor grid has COMPLETED and its writes are visible -- nothing that reads or writes global memory may come
// before it -- and pdl_trigger() lets this kernel's own successor start its prologue.  Both are no-ops for a
// kernel launched without the attribute.  Only kernels that call pdl_wait() may be launched through launch_k()
// (a kernel without the wait would simply run concurrently with its predecessor).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_wait();
  pdl_trigger();
}

// Kernel classes for the switch: tensor kernels (GEMM, attention: trigger once their last operand tile is
// requested) and row kernels (LayerNorm, column sums, casts: trigger right after their own wait).
enum : int { kPdlTensor = 1, kPdlRow = 2 };
int pdl_mode();  // runtime.cu: mb_set_pdl() / environment MB_PDL, bit mask of the classes above

template <int CLS = kPdlTensor, typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mode() & CLS) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_row(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  return launch_k<kPdlRow>(kern, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

// "do this once per device" (cudaFuncSetAttribute is a per-device setting): first() is true the first time
// it is called with a given current device.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return true;
    d &= 63;
    const bool was = done[d];
    done[d] = true;
    return !was;
  }
};

}  // namespace mb200
