// Re-sharding of an alpha-sharded CI vector between two row layouts: the all-to-all exchange step of the sharded engine
// (kernel K7 "distributed transpose"; no counterpart in the reference, which is single-process).
//
// Layout A groups the alpha strings (rows of C[Ia][Ib]) by the occupation of the FIRST log2(G) orbitals, layout B by the
// occupation of the LAST log2(G) orbitals.  A tUPS brick on the orbital pair (p, p+1) is purely local in layout A when
// p >= log2(G) and in layout B when p + 1 < n - log2(G), so a circuit runs as a few local phases separated by re-shards
// (slowquant_b200/distributed.py plans the phases).  A re-shard moves every row exactly once: the owner of a row in the
// source layout writes it straight into the destination owner's buffer through the CUDA-IPC peer mapping over
// NVLink / NVSwitch -- one kernel, no packing, no staging copies.  A row is a contiguous run of NB doubles, so the data
// moves with the bulk-copy engine (cp.async.bulk global -> shared -> peer global, mbarrier-tracked multi-stage pipeline
// driven by one thread per CTA: UBLKCP in the SASS) and never touches the LSU pipe; rows that are not 16-byte aligned
// (odd NB) take the plain vector load/store kernel instead.
#include <cstdio>
#include <cstdlib>

#include "sqsv_internal.h"

struct ReshardPeers {
  double* p[SQ_MAX_WORLD];
};

#define RESHARD_STAGE_BYTES 16384
#define RESHARD_STAGES 4
#define RESHARD_CHUNK_BYTES (8 * RESHARD_STAGE_BYTES)   // one CTA moves up to 128 KiB of one row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}

// One CTA = one chunk (<= RESHARD_CHUNK_BYTES) of one row.  Thread 0 drives the pipeline: RESHARD_STAGES bulk loads in
// flight, every landed stage leaves with a bulk store to the destination rank's buffer.
__global__ void __launch_bounds__(32)
reshard_bulk_kernel(const double* __restrict__ src, int64_t NB, const int32_t* __restrict__ dst_rank,
                    const int32_t* __restrict__ dst_row, const ReshardPeers peers, int chunks_per_row) {
  extern __shared__ __align__(128) unsigned char stage[];
  __shared__ __align__(8) uint64_t bars[RESHARD_STAGES];
  if (threadIdx.x != 0) return;
  const int64_t row = (int64_t)(blockIdx.x / chunks_per_row);
  const int chunk = (int)(blockIdx.x % chunks_per_row);
  const int64_t row_bytes = NB * 8;
  const int64_t off = (int64_t)chunk * RESHARD_CHUNK_BYTES;
  const int64_t left = row_bytes - off;
  const int nbytes = (int)(left < RESHARD_CHUNK_BYTES ? left : RESHARD_CHUNK_BYTES);   // multiple of 16 (NB even)
  const char* s = reinterpret_cast<const char*>(src + row * NB) + off;
  char* d = reinterpret_cast<char*>(peers.p[__ldg(dst_rank + row)] + (int64_t)__ldg(dst_row + row) * NB) + off;
  const uint32_t sbase = smem_u32(stage), bbase = smem_u32(bars);
  for (int i = 0; i < RESHARD_STAGES; ++i) mbar_init(bbase + 8u * i, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int n_st = (nbytes + RESHARD_STAGE_BYTES - 1) / RESHARD_STAGE_BYTES;
  auto load = [&](int i) {
    const int slot = i % RESHARD_STAGES;
    const int b = min(RESHARD_STAGE_BYTES, nbytes - i * RESHARD_STAGE_BYTES);
    mbar_expect_tx(bbase + 8u * slot, (uint32_t)b);
    bulk_g2s(sbase + (uint32_t)slot * RESHARD_STAGE_BYTES, s + (int64_t)i * RESHARD_STAGE_BYTES, (uint32_t)b, bbase + 8u * slot);
  };
  for (int i = 0; i < n_st && i < RESHARD_STAGES; ++i) load(i);
  for (int i = 0; i < n_st; ++i) {
    const int slot = i % RESHARD_STAGES;
    const int b = min(RESHARD_STAGE_BYTES, nbytes - i * RESHARD_STAGE_BYTES);
    mbar_wait(bbase + 8u * slot, (uint32_t)((i / RESHARD_STAGES) & 1));
    bulk_s2g(d + (int64_t)i * RESHARD_STAGE_BYTES, sbase + (uint32_t)slot * RESHARD_STAGE_BYTES, (uint32_t)b);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // the slot of the PREVIOUS store is free once that store has read its shared-memory source
    if (i >= 1 && i - 1 + RESHARD_STAGES < n_st) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      load(i - 1 + RESHARD_STAGES);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA retires
}

// Fallback for rows that are not 16-byte aligned: 8-byte loads / stores, one CTA per (row, 4096-double chunk).
__global__ void __launch_bounds__(256)
reshard_lsu_kernel(const double* __restrict__ src, int64_t NB, const int32_t* __restrict__ dst_rank,
                   const int32_t* __restrict__ dst_row, const ReshardPeers peers, int chunks_per_row, int vec2) {
  const int64_t row = (int64_t)(blockIdx.x / chunks_per_row);
  const int chunk = (int)(blockIdx.x % chunks_per_row);
  const int64_t c0 = (int64_t)chunk * 4096, c1 = min(c0 + 4096, NB);
  const double* s = src + row * NB;
  double* d = peers.p[__ldg(dst_rank + row)] + (int64_t)__ldg(dst_row + row) * NB;
  if (vec2) {   // NB even and 16-byte aligned bases
    const double2* s2 = reinterpret_cast<const double2*>(s + c0);
    double2* d2 = reinterpret_cast<double2*>(d + c0);
    const int n2 = (int)((c1 - c0) >> 1);
    for (int t = threadIdx.x; t < n2; t += 256) d2[t] = __ldcs(s2 + t);
  } else {
    for (int64_t t = c0 + threadIdx.x; t < c1; t += 256) d[t] = __ldcs(s + t);
  }
}

static int g_reshard_mode = 0;   // 0 bulk-copy engine when the rows are 16-byte aligned, 1 vector load/store kernel
void sq_reshard_set_mode(int lsu) { g_reshard_mode = lsu ? 1 : 0; }

// Move the n_rows local rows of `src_dev` (row r = NB doubles at src_dev + r * NB) to their owners in the other layout:
// row r goes to rank dst_rank_dev[r], local row dst_row_dev[r] of the buffer dst_ptrs_host[rank] (peer-mapped into this
// process; the own rank's entry is a local pointer).  Every destination row is written by exactly one source row, so no
// synchronisation is needed inside the kernel; the caller separates it from the neighbouring kernels of OTHER ranks with a
// device-wide barrier (all shards complete before / all rows landed after).
extern "C" int sq_reshard_rows(int device, int64_t n_rows, int64_t NB, const double* src_dev, const int32_t* dst_rank_dev,
                               const int32_t* dst_row_dev, double* const* dst_ptrs_host, int world, void* stream) {
  SqRange nvtx_range("sq_reshard_rows");
  if (n_rows < 0 || NB < 1 || world < 1 || world > SQ_MAX_WORLD || !dst_ptrs_host) return SQ_ERR_INVALID;
  if (n_rows == 0) return SQ_OK;
  if (!src_dev || !dst_rank_dev || !dst_row_dev) return SQ_ERR_INVALID;
  SQ_CUDA(cudaSetDevice(device));
  ReshardPeers peers;
  bool aligned = (NB % 2 == 0) && ((uintptr_t)src_dev % 16 == 0);
  for (int r = 0; r < SQ_MAX_WORLD; ++r) {
    peers.p[r] = r < world ? dst_ptrs_host[r] : nullptr;
    if (r < world && !peers.p[r]) {
      sq_set_error("sq_reshard_rows: missing destination pointer for rank %d", r);
      return SQ_ERR_INVALID;
    }
    if (r < world && (uintptr_t)peers.p[r] % 16 != 0) aligned = false;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  static const bool env_lsu = getenv("SQ_RESHARD_KERNEL") && getenv("SQ_RESHARD_KERNEL")[0] == 'l';   // A/B runs of whole programs
  if (aligned && !g_reshard_mode && !env_lsu) {
    const int64_t row_bytes = NB * 8;
    const int cpr = (int)((row_bytes + RESHARD_CHUNK_BYTES - 1) / RESHARD_CHUNK_BYTES);
    const int64_t n_cta = n_rows * cpr;
    if (n_cta > 0x7fffffffLL) {
      sq_set_error("sq_reshard_rows: %lld chunks exceed the grid limit", (long long)n_cta);
      return SQ_ERR_INVALID;
    }
    const int smem = RESHARD_STAGES * RESHARD_STAGE_BYTES;
    static bool attr = false;
    if (!attr) {
      e = cudaFuncSetAttribute(reshard_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr = (e == cudaSuccess);
    }
    if (e == cudaSuccess) {
      reshard_bulk_kernel<<<(unsigned)n_cta, 32, smem, st>>>(src_dev, NB, dst_rank_dev, dst_row_dev, peers, cpr);
      e = cudaGetLastError();
    }
  } else {
    const int cpr = (int)((NB + 4095) / 4096);
    const int64_t n_cta = n_rows * cpr;
    if (n_cta > 0x7fffffffLL) {
      sq_set_error("sq_reshard_rows: %lld chunks exceed the grid limit", (long long)n_cta);
      return SQ_ERR_INVALID;
    }
    reshard_lsu_kernel<<<(unsigned)n_cta, 256, 0, st>>>(src_dev, NB, dst_rank_dev, dst_row_dev, peers, cpr, aligned ? 1 : 0);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    sq_set_error("reshard kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}
