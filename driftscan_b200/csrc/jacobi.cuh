// Batched block one-sided Jacobi (svd.cu), shared by the SVD chain and the KL eigensolver.
#pragma once
#include "dsb_common.cuh"

namespace dsb {

typedef cplx<double> zc;

struct JacobiScratch {
  unsigned long long *amax = nullptr;                        // [batch]
  int32_t *rot = nullptr, *done = nullptr, *flag = nullptr;  // [batch], [batch], [2]
  int32_t *skip = nullptr;                                   // [batch][npairs_ld]
  zc *W = nullptr;                                           // [batch][npairs_ld][32 * 32]
  int npairs_ld = 0;
  double *nrm2 = nullptr;       // [batch][ldr_max] row norms by position, refreshed every sweep
  int32_t *nlive = nullptr;     // [batch] live row blocks
  int32_t *blkstamp = nullptr;  // [batch][nblk_ld] step at which a block was last modified
  int32_t *pairstamp = nullptr; // [batch][nblk_ld][nblk_ld] step at which a pair was last found orthogonal
  int nblk_ld = 0, ldr_max = 0;
  int32_t *h_flag = nullptr;  // pinned [2]
  int alloc(int batch, int nrows_max, cudaStream_t stream);
  void release(cudaStream_t stream);
};

// One Jacobi pass over the active rows idx[b][0..nact[b]) of every matrix R[b] (ldr x ncols, row
// major): the rows become mutually orthogonal with respect to columns [ip0, ip1); every column
// is carried along.  sweeps[b] (device) receives the sweep count; `nmax` = upper bound of nact
// (-1: read it back from the device).
int jacobi_pass(zc *R, int ldr, int ncols, const int32_t *idx, const int32_t *nact, int batch, int ip0, int ip1,
                int nmax, int max_sweeps, double tol, int32_t *sweeps, JacobiScratch &js, cudaStream_t stream);

// Unitary row operations (Householder reflections) that bring columns [ip0, ip1) of the active
// rows to upper-triangular form before a Jacobi pass: cuts the sweep count of graded matrices
// from ~25 to ~5.  Every column is carried along, exactly as in jacobi_pass.
// `keep` != NULL: the reflectors are handed to the caller (stream-ordered allocations it frees with
// cudaFreeAsync) instead of being released: H_k = I - tau_k v_k v_k^H, v_k = Vh[b][k][position],
// positions = the order of idx at the time of the call.
struct HHKeep {
  zc *Vh = nullptr;        // [batch][steps][ldr]
  double *tauh = nullptr;  // [batch][steps]
  int steps = 0, nmax = 0;
};
int householder_precondition(zc *R, int ldr, int ncols, const int32_t *idx, const int32_t *nact, int batch, int ip0,
                             int ip1, int nmax, JacobiScratch &js, cudaStream_t stream, HHKeep *keep = nullptr);

}  // namespace dsb
