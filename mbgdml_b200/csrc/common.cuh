// Shared definitions for the mbgdml_b200 CUDA translation unit (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/mbgdml_b200.h"

namespace mbg {

constexpr int kMaxOrder = MBG_MAX_ORDER;
constexpr int kMaxFragAtoms = MBG_MAX_FRAG_ATOMS;
constexpr int kMaxAlchemy = MBG_MAX_ALCHEMY;
constexpr int kCullThreads = 256;  // candidates per cull/compact block
constexpr int kShardGranule = 16;  // sharding granule: consecutive candidates that stay on one rank (divides kCullThreads); 64 until r2: 8 ranks of box140 x 64 frames were 0.56 % apart

// One periodic cell on device (mbgdml/periodic.py:32-140): lives in device memory so that every frame of a batch can
// have its own (NPT trajectories, the strained boxes of the finite-difference virial, stress.py:75-152) and so that
// a captured CUDA graph of a step sees a new cell without being re-captured.
struct CellDev {
  double cell[9], cell_inv[9];    // row vectors; f = v @ cell_inv
  double rcell[9], rcell_inv[9];  // Minkowski-reduced cell for the general branch
  double half_min_len;            // 0.5 * min |cell vector|
  double cutoff;                  // Cell.cutoff
  int fast_reject;                // orthorhombic cell and cutoff <= half_min_len: a naive image beyond L_min/2 implies rejection
  int ortho;                      // orthorhombic, fully periodic cell: the naive image is the exact minimum image (pairlist.cuh)
  int pbc[3];                     // 1: periodic along this cell vector (Cell.pbc, periodic.py:40-59).  With a 0 entry find_mic
                                  // always takes the general branch: half_min_len = -1 makes every "naive image is safe" test fail
  int pad;
};

// The supersystem on device.
struct SysDev {
  int n_atoms;
  int n_entities;
  const int* ent_off;    // [n_entities+1]
  const int* ent_atoms;  // CSR values
  const double* masses;  // [n_atoms]
  int periodic;          // 0: cluster, 1: minimum-image under cells[frame * cell_stride]
  int cell_stride;       // 0: one cell for all frames, 1: one per frame
  const CellDev* cells;
};

__device__ __forceinline__ const CellDev& cell_of(const SysDev& s, long long frame) { return s.cells[frame * s.cell_stride]; }

// Acceptance criteria of one model.
struct CritDev {
  int kind, bound;
  double lo, hi;
  int n_groups;
  int group[kMaxFragAtoms];  // entity label (0..n_groups-1) of each fragment atom, for com_distance_sum
};

// Fragment source + alchemy of one job.
struct JobDev {
  int order;
  int n_frag_atoms;
  int explicit_combs;       // 0: product of sets (ascending filter), 1: explicit tuples
  int homogeneous;          // 1: every slot draws from the same ascending set -> candidates are the C(n, order)
                            //    combinations, unranked in lexicographic order (== product order of the ascending tuples)
  int n_set;                // size of that set
  const long long* binom;   // [(n_set + 1) x (order + 1)] binomial table, binom[m * (order + 1) + r] = C(m, r)
  int set_off[kMaxOrder + 1];
  int slot_atoms[kMaxOrder]; // atoms per entity of each slot (sets only)
  int sets_ascending;        // every slot's set lists its entities in ascending order (what get_avail_entities yields)
  const int* set_ent;       // concatenated per-slot entity ids (device)
  const int* combs;         // [n][order] (device) when explicit
  long long cand_per_frame; // product size, C(n_set, order), or n_combs
  int n_alch;
  int alch_ent[kMaxAlchemy];
  double alch_fac[kMaxAlchemy];
};

}  // namespace mbg
