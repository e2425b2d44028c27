// K1 for large systems: pair-list candidate generation.
//
// k_cull (enumerate.cuh) unranks and tests every one of the C(n, order) tuples of a frame -- the reference does the
// same (utils.gen_combs feeds all of them to the slice -> MIC -> criteria loop, gdml_predict.py:210-235) -- so its
// cost grows with n^3 while the accepted fragments grow with n.  Here a pre-pass builds, per frame, the bit matrix of
// entity pairs (a < b) that CAN be together in an accepted fragment, and the candidates that are evaluated in full are
// only the tuples (a, b, c) with b, c both neighbours of a and c a neighbour of b.  The pair test is a NECESSARY
// condition of acceptance, decided only beyond the 1e-9 guard band (geom.cuh), so the full evaluation
// (first_atoms_may_pass / fragment_accept, unchanged) still takes every decision that counts and the accepted set,
// its order (frame, then lexicographic tuple == candidate rank) and the fragment records are identical to k_cull's.
//
// Job shapes: one homogeneous entity set, 2- to 6-body (a row's reduced candidates are the C(deg, order - 1)
// combinations of its neighbours, every pair among them checked against the bits), and heterogeneous product spaces of
// 2- and 3-body jobs ([h2o, meoh], [h2o, h2o, meoh], ...: one bit row per later slot, the ascending filter of
// utils.gen_combs folded into the bits).  A plain call reads the reduced total back once when the job is large
// (api.cu: kPairListSyncCand); otherwise, and inside captured steps, every kernel reads it on the device.
//
// Pair conditions (p_x = first atom of entity x, r = the fragment's final coordinates):
//   * periodic cell (periodic.py:61-107): every r_i - r_j is p_i - p_j plus a lattice vector, so
//     |r_i - r_j| >= the minimum-image length of p_i - p_j, and periodic.py:80-83 rejects when any pair distance
//     reaches the cut-off  ->  need mic(p_a - p_b) < cutoff (and < the bound of max_atom_pair_dist, if that is the
//     criterion).  The minimum image is the naive one in an orthorhombic cell, the Minkowski-reduced 27-image search
//     otherwise (both exact minima).
//   * cluster + com_distance_sum < hi (descriptors.py:159-201): sum_g |CM - CM_g| >= |CM - CM_a| + |CM - CM_b|
//     >= |CM_a - CM_b|  ->  need |CM_a - CM_b| < hi.
//   * cluster + max_atom_pair_dist < hi (descriptors.py:131-156): need |p_a - p_b| < hi.
#pragma once
#include "enumerate.cuh"

namespace mbg {

constexpr int kPairWarps = 8;  // rows (frame, a) per block of k_pair_bits

struct PairListDev {
  unsigned long long* adj;  // [n_frames][n_set][W]: bit b of row a set <=> b > a and the pair may pass
  int* deg;                 // [n_frames * n_set]: set bits of a row
  long long* row_cnt;       // [n_frames * n_set]: reduced candidates of a row: C(deg, order - 1)
  long long* row_off;       // [n_frames * n_set + 1]: exclusive prefix sum of row_cnt; the last entry is the total
  int W;                    // 64-bit words per row
  int n_rows;               // n_frames * n0
  int n0;                   // rows per frame: entities of slot 0 (== n_set for a homogeneous job)
  // Heterogeneous product spaces ([h2o, h2o, meoh], ...): a row (frame, i0) has one bit set per slot -- `adj` over the
  // entities of slot 1, `adj2` over those of slot 2 -- with bit p set <=> entity p of that slot is above the row's
  // entity (the ascending filter of utils.gen_combs, utils.py:479-488) and the pair may pass; the row's reduced
  // candidates are the deg x deg2 combinations in product order.
  int hetero;
  unsigned long long* adj2;  // [n_rows][W2]
  int* deg2;                 // [n_rows]
  int W2;
  int mode;                 // 0: periodic cut-off, 1: cluster centre-of-mass distance, 2: cluster first-atom distance
  double thr;               // modes 1, 2: the criterion's upper bound; mode 0: its bound if max_atom_pair_dist, else +inf
};

__device__ __forceinline__ bool pair_may_pass(const SysDev& sys, const PairListDev& nl, long long frame,
                                              const double* __restrict__ Rf, int ea, int eb) {
  const int a0 = sys.ent_off[ea], b0 = sys.ent_off[eb];
  if (nl.mode == 1) {
    double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0}, ma = 0, mb = 0;
    for (int i = a0; i < sys.ent_off[ea + 1]; ++i) {
      const int at = sys.ent_atoms[i];
      const double m = sys.masses[at];
      ca[0] += m * Rf[3ll * at], ca[1] += m * Rf[3ll * at + 1], ca[2] += m * Rf[3ll * at + 2], ma += m;
    }
    for (int i = b0; i < sys.ent_off[eb + 1]; ++i) {
      const int at = sys.ent_atoms[i];
      const double m = sys.masses[at];
      cb[0] += m * Rf[3ll * at], cb[1] += m * Rf[3ll * at + 1], cb[2] += m * Rf[3ll * at + 2], mb += m;
    }
    const double len = norm3_rn(ca[0] / ma - cb[0] / mb, ca[1] / ma - cb[1] / mb, ca[2] / ma - cb[2] / mb);
    return !(len >= nl.thr * kGuard);
  }
  const double* pa = Rf + 3ll * sys.ent_atoms[a0];
  const double* pb = Rf + 3ll * sys.ent_atoms[b0];
  const double d[3] = {__dsub_rn(pb[0], pa[0]), __dsub_rn(pb[1], pa[1]), __dsub_rn(pb[2], pa[2])};
  if (nl.mode == 2) return !(norm3_rn(d[0], d[1], d[2]) >= nl.thr * kGuard);
  const CellDev& s = cell_of(sys, frame);
  double v[3], len;
  if (s.ortho) {
    len = mic_naive(s, d, v);
  } else {
    mic_general(s, d, v);
    len = norm3_rn(v[0], v[1], v[2]);
  }
  const double thr = s.cutoff < nl.thr ? s.cutoff : nl.thr;
  return !(len >= thr * kGuard);
}

// One warp per row (frame, a): lanes take the 64-bit words of the row in turn.
__global__ void __launch_bounds__(kPairWarps * 32) k_pair_bits(const SysDev sys, const JobDev job, const PairListDev nl,
                                                               const double* __restrict__ R) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kPairWarps + (threadIdx.x >> 5);
  if (row >= nl.n_rows) return;
  const int n = nl.n0;
  const long long frame = row / n;
  const int a = (int)(row - frame * n);
  const double* Rf = R + frame * 3ll * sys.n_atoms;
  const int ea = job.set_ent[a];
  if (nl.hetero) {
    int degs[2] = {0, 0};
    for (int s = 1; s < job.order; ++s) {
      const int* set = job.set_ent + job.set_off[s];
      const int ns = job.set_off[s + 1] - job.set_off[s], Ws = s == 1 ? nl.W : nl.W2;
      unsigned long long* out = s == 1 ? nl.adj + row * nl.W : nl.adj2 + row * nl.W2;
      int d = 0;
      for (int w = lane; w < Ws; w += 32) {
        unsigned long long bits = 0;
        const int p_hi = min(w * 64 + 64, ns);
        for (int p = w * 64; p < p_hi; ++p) {
          const int eb = set[p];
          if (eb > ea && pair_may_pass(sys, nl, frame, Rf, ea, eb)) bits |= 1ull << (p & 63);
        }
        out[w] = bits;
        d += __popcll(bits);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      degs[s - 1] = d;
    }
    if (lane == 0) {
      nl.deg[row] = degs[0];
      if (job.order == 3) nl.deg2[row] = degs[1];
      nl.row_cnt[row] = job.order == 3 ? (long long)degs[0] * degs[1] : degs[0];
    }
    return;
  }
  int deg = 0;
  for (int w = lane; w < nl.W; w += 32) {
    unsigned long long bits = 0;
    const int b_lo = max(w * 64, a + 1), b_hi = min(w * 64 + 64, n);
    for (int b = b_lo; b < b_hi; ++b)
      if (pair_may_pass(sys, nl, frame, Rf, ea, job.set_ent[b])) bits |= 1ull << (b & 63);
    nl.adj[row * nl.W + w] = bits;
    deg += __popcll(bits);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) deg += __shfl_xor_sync(0xffffffffu, deg, o);
  if (lane == 0) {
    nl.deg[row] = deg;
    // C(deg, order - 1) combinations of the row's neighbours (orders above 3: from the job's binomial table)
    nl.row_cnt[row] = job.order == 2 ? deg : job.order == 3 ? (long long)deg * (deg - 1) / 2
                                                            : job.binom[(long long)deg * (job.order + 1) + (job.order - 1)];
  }
}

// Exclusive prefix sum of n 64-bit counts by one CTA; off[n] = total.
__global__ void __launch_bounds__(1024) k_scan_rows(const long long* __restrict__ cnt, int n, long long* __restrict__ off) {
  __shared__ long long warp_sums[32];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = i < n ? cnt[i] : 0;
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
      warp_sums[lane] = w;
    }
    __syncthreads();
    const long long before = carry + (warp ? warp_sums[warp - 1] : 0) + (x - v);
    if (i < n) off[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = carry;
}

// Position of the k-th (0-based) set bit of a 64-bit word (k < popcount): halving search on popcounts.
__host__ __device__ __forceinline__ int select64(unsigned long long x, int k) {
  int pos = 0;
#pragma unroll
  for (int width = 32; width >= 1; width >>= 1) {
    const unsigned long long low = x & ((1ull << width) - 1ull);
#ifdef __CUDA_ARCH__
    const int c = __popcll(low);
#else
    const int c = __builtin_popcountll(low);
#endif
    if (k >= c) k -= c, x >>= width, pos += width;
    else x = low;
  }
  return pos;
}

// Position of the k-th (0-based) set bit of a row, counted from its first word.
__device__ __forceinline__ int select_bit(const unsigned long long* __restrict__ rowbits, int W, int k) {
  for (int w = 0; w < W; ++w) {
    const unsigned long long x = rowbits[w];
    const int c = __popcll(x);
    if (k < c) return w * 64 + select64(x, k);
    k -= c;
  }
  return -1;
}

// Reduced candidate g -> (frame, positions in the slots' sets, entity tuple).  Returns 1 when the tuple goes to the full
// evaluation, 2 when its last pair is not admissible (order 3: (b, c) fails the pair test), i.e. the tuple cannot be
// accepted, 0 when it is not ascending (heterogeneous jobs only: utils.gen_combs never yields it).
__device__ __forceinline__ int decode_reduced(const SysDev& sys, const double* __restrict__ R, const PairListDev& nl,
                                              const JobDev& j, long long g, long long* frame, int pos[kMaxOrder],
                                              int ent[kMaxOrder]) {
  int lo = 0, hi = nl.n_rows - 1;  // last row with row_off[row] <= g
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (nl.row_off[mid] <= g) lo = mid; else hi = mid - 1;
  }
  const long long row = lo, local = g - nl.row_off[row];
  const int n = nl.n0;
  *frame = row / n;
  pos[0] = (int)(row - *frame * n);
  const unsigned long long* bits = nl.adj + row * nl.W;
  if (nl.hetero) {
    ent[0] = j.set_ent[pos[0]];
    if (j.order == 2) {
      pos[1] = select_bit(bits, nl.W, (int)local);
      ent[1] = j.set_ent[j.set_off[1] + pos[1]];
      return 1;
    }
    const long long d2 = nl.deg2[row];
    pos[1] = select_bit(bits, nl.W, (int)(local / d2));
    pos[2] = select_bit(nl.adj2 + row * nl.W2, nl.W2, (int)(local % d2));
    ent[1] = j.set_ent[j.set_off[1] + pos[1]];
    ent[2] = j.set_ent[j.set_off[2] + pos[2]];
    if (!(ent[1] < ent[2])) return 0;
    if (R == nullptr) return 1;  // compaction / mask of an accepted candidate: only the tuple is wanted
    return pair_may_pass(sys, nl, *frame, R + *frame * 3ll * sys.n_atoms, ent[1], ent[2]) ? 1 : 2;
  }
  bool ok = true;
  if (j.order == 2) {
    pos[1] = select_bit(bits, nl.W, (int)local);
  } else if (j.order > 3) {
    // (order - 1)-subsets of the row's deg neighbours in lexicographic order: decode_tuple's unranking with n = deg
    // (the job's binomial table covers every m <= n_set), then all pairs among them must be in the list
    const int n_nb = nl.deg[row], m = j.order - 1, w = j.order + 1;
    long long rem = local;
    int prev = -1;
    for (int s = 0; s < m; ++s) {
      const int r = m - s - 1;
      int p;
      if (r == 0) {
        p = prev + 1 + (int)rem;
      } else {
        const long long top = j.binom[(long long)(n_nb - prev - 1) * w + (r + 1)];
        const long long target = top - rem;
        int lo2 = r, hi2 = n_nb - prev - 2;
        while (lo2 < hi2) {
          const int mid = (lo2 + hi2 + 1) >> 1;
          if (j.binom[(long long)mid * w + (r + 1)] < target) lo2 = mid; else hi2 = mid - 1;
        }
        p = n_nb - 1 - lo2;
        rem -= top - j.binom[(long long)(n_nb - p) * w + (r + 1)];
      }
      pos[s + 1] = select_bit(bits, nl.W, p);
      prev = p;
    }
    for (int s = 1; s < j.order && ok; ++s) {
      const unsigned long long* bits_s = nl.adj + (*frame * n + pos[s]) * nl.W;
      for (int t = s + 1; t < j.order; ++t) ok = ok && ((bits_s[pos[t] >> 6] >> (pos[t] & 63)) & 1ull);
    }
  } else {
    // pairs (i < k) of the row's deg neighbours in lexicographic order: S(i) = i deg - i (i + 1) / 2 pairs start below i
    const long long deg = nl.deg[row];
    const double t = (double)(2 * deg - 1);
    long long i = (long long)((t - sqrt(t * t - 8.0 * (double)local)) * 0.5);
    i = i < 0 ? 0 : (i > deg - 2 ? deg - 2 : i);
    while (i > 0 && i * deg - i * (i + 1) / 2 > local) --i;
    while ((i + 1) * deg - (i + 1) * (i + 2) / 2 <= local) ++i;
    const long long k = local - (i * deg - i * (i + 1) / 2) + i + 1;
    pos[1] = select_bit(bits, nl.W, (int)i);
    pos[2] = select_bit(bits, nl.W, (int)k);
    const unsigned long long* bits_b = nl.adj + (*frame * n + pos[1]) * nl.W;
    ok = (bits_b[pos[2] >> 6] >> (pos[2] & 63)) & 1ull;
  }
  for (int s = 0; s < j.order; ++s) ent[s] = j.set_ent[pos[s]];
  return ok ? 1 : 2;
}

// Lexicographic rank of the combination pos[0] < pos[1] < ... among the C(n_set, order) (inverse of decode_tuple's
// unranking): the candidate index within the frame, i.e. the row of the decomposed outputs.
__device__ __forceinline__ long long rank_combination(const JobDev& j, const int pos[kMaxOrder]) {
  if (!j.homogeneous) {  // product order of the slots' sets, last slot fastest (decode_tuple)
    long long idx = 0;
    for (int s = 0; s < j.order; ++s) idx = idx * (j.set_off[s + 1] - j.set_off[s]) + pos[s];
    return idx;
  }
  const int n = j.n_set, w = j.order + 1;
  long long rank = 0;
  int prev = -1;
  for (int s = 0; s < j.order; ++s) {
    const int r = j.order - s - 1, p = pos[s];
    if (r == 0) rank += p - prev - 1;
    else rank += j.binom[(long long)(n - prev - 1) * w + (r + 1)] - j.binom[(long long)(n - p) * w + (r + 1)];
    prev = p;
  }
  return rank;
}

// k_cull over the reduced candidate space, sharded like the full space.  The reduced total is read from the device:
// a launch inside a captured step is sized for the full space (a.n_cand_total is only that upper bound there) and the
// blocks beyond the reduced total flag nothing.
template <int NA>
__global__ void __launch_bounds__(kCullThreads) k_cull_nl(const __grid_constant__ CullArgs a, const PairListDev nl) {
  const long long g = global_candidate(a, blockIdx.x, threadIdx.x);
  int flag = 0;
  if (g < nl.row_off[nl.n_rows]) {
    long long frame;
    int pos[kMaxOrder], ent[kMaxOrder];
    flag = decode_reduced(a.sys, a.R, nl, a.job, g, &frame, pos, ent);
    if (flag == 1) {
      flag = 2;
      const double* Rf = a.R + frame * 3ll * a.sys.n_atoms;
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
  if (threadIdx.x == 0) {
    a.block_accept[blockIdx.x] = n_acc;
    a.block_enum[blockIdx.x] = 0;  // the enumerated count comes from k_add_total / k_count_ascending
  }
}

__global__ void __launch_bounds__(kCullThreads) k_compact_nl(const __grid_constant__ CompactArgs a, const PairListDev nl) {
  __shared__ int warp_cnt[kCullThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool acc = a.flags[(long long)blockIdx.x * kCullThreads + tid] == 1;
  const unsigned ballot = __ballot_sync(0xffffffffu, acc);
  if (lane == 0) warp_cnt[warp] = __popc(ballot);
  __syncthreads();
  if (!acc) return;
  int rank = __popc(ballot & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) rank += warp_cnt[w];
  const long long pos_out = a.rec_base + a.block_offsets[blockIdx.x] + rank;
  CullArgs ca;  // only the fields global_candidate reads
  ca.granule0 = a.granule0, ca.shard_count = a.shard_count, ca.shard_rank = a.shard_rank;
  const long long g = global_candidate(ca, blockIdx.x, tid);
  long long frame;
  int pos[kMaxOrder], ent[kMaxOrder], atoms[kMaxFragAtoms];
  decode_reduced(a.sys, nullptr, nl, a.job, g, &frame, pos, ent);
  const int na = gather_atoms(a.sys, a.job, ent, atoms);
  a.rec_frame[pos_out] = (int)frame;
  a.rec_slot[pos_out] = rank_combination(a.job, pos);
  a.rec_lambda[pos_out] = alchemy_lambda(a.job, ent);
  for (int k = 0; k < na; ++k) a.rec_atoms[pos_out * a.job.n_frag_atoms + k] = atoms[k];
}

// mbg_accept_mask through the pair list: accepted reduced candidates set their byte of the (zeroed) full-space mask.
__global__ void __launch_bounds__(kCullThreads) k_mask_nl(const __grid_constant__ CullArgs a, const PairListDev nl,
                                                          unsigned char* __restrict__ mask, long long mask_lo, long long mask_n) {
  if (a.flags[(long long)blockIdx.x * kCullThreads + threadIdx.x] != 1) return;
  const long long g = global_candidate(a, blockIdx.x, threadIdx.x);
  long long frame;
  int pos[kMaxOrder], ent[kMaxOrder];
  decode_reduced(a.sys, nullptr, nl, a.job, g, &frame, pos, ent);
  const long long idx = frame * a.job.cand_per_frame + rank_combination(a.job, pos) - mask_lo;
  if (idx >= 0 && idx < mask_n) mask[idx] = 1;
}

__global__ void k_add_total(long long* p, long long v) { *p += v; }

// Enumerated count of a heterogeneous job (the ascending tuples of its product space, per frame): one thread per
// (i0, i1) with ent0 < ent1 adds the entities of slot 2 above ent1 (sets are ascending: binary search).
__global__ void __launch_bounds__(256) k_count_ascending(const JobDev job, unsigned long long* __restrict__ count) {
  const int n0 = job.set_off[1] - job.set_off[0], n1 = job.set_off[2] - job.set_off[1];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long c = 0;
  if (t < (long long)n0 * n1) {
    const int e0 = job.set_ent[t / n1], e1 = job.set_ent[job.set_off[1] + t % n1];
    if (e0 < e1) {
      if (job.order == 2) {
        c = 1;
      } else {
        const int* s2 = job.set_ent + job.set_off[2];
        int lo = 0, hi = job.set_off[3] - job.set_off[2];  // first position with s2[p] > e1
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s2[mid] > e1) hi = mid; else lo = mid + 1;
        }
        c = (unsigned long long)(job.set_off[3] - job.set_off[2] - lo);
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// ... times the frames, this shard's share (the pair list shards the reduced space, so the enumerated tuples are
// attributed evenly: the sum over the shards is the job's count).
__global__ void k_add_share(long long* p, const unsigned long long* per_frame, long long n_frames, int rank, int count) {
  const long long total = (long long)*per_frame * n_frames;
  *p += total / count + (rank < total % count ? 1 : 0);
}

// mbg_accept_mask through the pair list, heterogeneous jobs: 2 where the tuple is ascending, else 0.
__global__ void __launch_bounds__(kCullThreads) k_mask_init(const JobDev job, long long n_total, unsigned char* __restrict__ mask) {
  const long long g = (long long)blockIdx.x * kCullThreads + threadIdx.x;
  if (g >= n_total) return;
  int ent[kMaxOrder];
  mask[g] = decode_tuple(job, g % job.cand_per_frame, ent) ? 2 : 0;
}

}  // namespace mbg
