// K1/K2: on-device fragment enumeration, culling and ordered compaction.
//
// Candidate space of one job = n_frames x (Cartesian product of the per-slot entity
// sets | explicit tuples).  A candidate survives when its tuple is strictly ascending
// (utils.gen_combs, reference mbgdml/utils.py:449-489), its minimum-image pair
// distances are below the cell cut-off (periodic.py:61-107) and the model's criteria
// accept it (descriptors.py:65-103).  Survivors are compacted IN CANDIDATE ORDER
// (== the reference's emission order) into fragment records for the contraction kernel.
//
// Sharding: the global candidate space is cut into granules of kShardGranule candidates;
// shard r of G owns granules r, r+G, r+2G, ...  (block-cyclic; no communication).  A cull block of
// kCullThreads threads covers kCullThreads / kShardGranule consecutive granules of its shard.
#pragma once
#include "geom.cuh"

namespace mbg {

struct CullArgs {
  SysDev sys;
  CritDev crit;
  JobDev job;
  const double* R;        // [n_frames][n_atoms][3]
  long long n_cand_total; // n_frames * cand_per_frame
  long long granule0;     // first block (kCullThreads candidates of this shard's own space) of this launch
  int shard_rank, shard_count;
  unsigned char* flags;   // [launch candidates] 0: dropped, 1: accepted, 2: ascending but rejected
  int* block_accept;      // [blocks] accepted per block
  int* block_enum;        // [blocks] ascending tuples per block
};

__device__ __forceinline__ long long global_candidate(const CullArgs& a, int block, int tid) {
  const long long local = (a.granule0 + block) * (long long)kCullThreads + tid;  // index in this shard's own space
  const long long granule = (local / kShardGranule) * a.shard_count + a.shard_rank;
  return granule * kShardGranule + local % kShardGranule;
}

// One thread per candidate: decode -> gather -> MIC/cut-off -> criteria -> flag.
template <int NA>
__global__ void __launch_bounds__(kCullThreads) k_cull(const __grid_constant__ CullArgs a) {
  const long long g = global_candidate(a, blockIdx.x, threadIdx.x);
  int flag = 0;
  if (g < a.n_cand_total) {
    const long long frame = g / a.job.cand_per_frame;
    const long long cand = g - frame * a.job.cand_per_frame;
    int ent[kMaxOrder];
    if (decode_tuple(a.job, cand, ent)) {
      const double* Rf = a.R + frame * 3ll * a.sys.n_atoms;
      flag = 2;
      if (first_atoms_may_pass(a.sys, frame, a.job, Rf, ent)) {
        int atoms[kMaxFragAtoms];
        gather_atoms(a.sys, a.job, ent, atoms);
        double r[NA][3];
        flag = fragment_accept<NA>(a.sys, frame, a.crit, Rf, atoms, r) ? 1 : 2;
      }
    }
  }
  a.flags[(long long)blockIdx.x * kCullThreads + threadIdx.x] = (unsigned char)flag;
  const int n_acc = __syncthreads_count(flag == 1);
  const int n_enum = __syncthreads_count(flag != 0);
  if (threadIdx.x == 0) {
    a.block_accept[blockIdx.x] = n_acc;
    a.block_enum[blockIdx.x] = n_enum;
  }
}

// Exclusive scan of per-block counts by one CTA (block counts are few: candidates/256).
// out_offsets[i] = sum_{k<i} counts[k]; totals[0] = sum counts, totals[1] = sum counts2.
// job_totals (optional): running (accepted, enumerated) sums of a job whose launches are chunked.
__global__ void __launch_bounds__(1024) k_scan_blocks(const int* __restrict__ counts, const int* __restrict__ counts2,
                                                      int n, long long* __restrict__ offsets,
                                                      long long* __restrict__ totals, long long* job_totals = nullptr) {
  __shared__ long long warp_sums[32];
  __shared__ long long carry, carry2;
  if (threadIdx.x == 0) carry = 0, carry2 = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long acc2 = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = i < n ? counts[i] : 0;
    if (i < n && counts2) acc2 += counts2[i];
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    const long long before = carry + (warp ? warp_sums[warp - 1] : 0) + (x - v);
    if (i < n) offsets[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  // reduce acc2
#pragma unroll
  for (int o = 16; o; o >>= 1) acc2 += __shfl_down_sync(0xffffffffu, acc2, o);
  if (lane == 0) warp_sums[warp] = acc2;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t2 = 0;
    for (int w = 0; w < 32; ++w) t2 += warp_sums[w];
    totals[0] = carry;
    totals[1] = t2;
    if (job_totals) job_totals[0] += carry, job_totals[1] += t2;
  }
}

struct CompactArgs {
  SysDev sys;
  JobDev job;
  long long n_cand_total;
  long long granule0;
  int shard_rank, shard_count;
  const unsigned char* flags;
  const long long* block_offsets;  // exclusive scan of block_accept
  long long rec_base;              // records already produced by earlier launches of this job
  // outputs (fragment records, in candidate order)
  int* rec_frame;       // [n_rec]
  int* rec_atoms;       // [n_rec][n_frag_atoms]
  double* rec_lambda;   // [n_rec]
  long long* rec_slot;  // [n_rec] candidate index within the frame (row of the decomposed outputs)
};

__global__ void __launch_bounds__(kCullThreads) k_compact(const __grid_constant__ CompactArgs a) {
  __shared__ int warp_cnt[kCullThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool acc = a.flags[(long long)blockIdx.x * kCullThreads + tid] == 1;
  const unsigned ballot = __ballot_sync(0xffffffffu, acc);
  if (lane == 0) warp_cnt[warp] = __popc(ballot);
  __syncthreads();
  if (!acc) return;
  int rank = __popc(ballot & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) rank += warp_cnt[w];
  const long long pos = a.rec_base + a.block_offsets[blockIdx.x] + rank;

  CullArgs ca;  // only the fields global_candidate reads
  ca.granule0 = a.granule0, ca.shard_count = a.shard_count, ca.shard_rank = a.shard_rank;
  const long long g = global_candidate(ca, blockIdx.x, tid);
  const long long frame = g / a.job.cand_per_frame;
  const long long cand = g - frame * a.job.cand_per_frame;
  int ent[kMaxOrder], atoms[kMaxFragAtoms];
  decode_tuple(a.job, cand, ent);
  const int na = gather_atoms(a.sys, a.job, ent, atoms);
  a.rec_frame[pos] = (int)frame;
  a.rec_slot[pos] = cand;
  a.rec_lambda[pos] = alchemy_lambda(a.job, ent);
  for (int k = 0; k < na; ++k) a.rec_atoms[pos * a.job.n_frag_atoms + k] = atoms[k];
}

// utils.gen_combs on device (mbg_enumerate): flag ascending tuples of the product space ...
__global__ void __launch_bounds__(kCullThreads) k_flag_ascending(const JobDev job, long long cand0, long long n_cand,
                                                                 unsigned char* flags, int* block_counts) {
  const long long c = cand0 + (long long)blockIdx.x * kCullThreads + threadIdx.x;
  int ent[kMaxOrder];
  const bool asc = (c < n_cand) && decode_tuple(job, c, ent);
  flags[(long long)blockIdx.x * kCullThreads + threadIdx.x] = asc ? 1 : 0;
  const int n = __syncthreads_count(asc);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = n;
}

// ... and write them out in product order.
__global__ void __launch_bounds__(kCullThreads) k_write_combs(const JobDev job, long long cand0,
                                                              const unsigned char* flags,
                                                              const long long* block_offsets, long long base,
                                                              int* combs_out) {
  __shared__ int warp_cnt[kCullThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool acc = flags[(long long)blockIdx.x * kCullThreads + tid] == 1;
  const unsigned ballot = __ballot_sync(0xffffffffu, acc);
  if (lane == 0) warp_cnt[warp] = __popc(ballot);
  __syncthreads();
  if (!acc) return;
  int rank = __popc(ballot & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) rank += warp_cnt[w];
  const long long pos = base + block_offsets[blockIdx.x] + rank;
  int ent[kMaxOrder];
  decode_tuple(job, cand0 + (long long)blockIdx.x * kCullThreads + tid, ent);
  for (int s = 0; s < job.order; ++s) combs_out[pos * job.order + s] = ent[s];
}

}  // namespace mbg
