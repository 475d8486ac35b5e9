// visual_api.cuh — host side of boundary B2 (include/lvba_b200.h): problem set-up mirroring the Ceres
// problem built at reference src/lvba_system.cpp:1578-1640 and the trust-region LM loop of
// ceres-solver 2.1.0 (TrustRegionMinimizer + LevenbergMarquardtStrategy).
#pragma once
#include <cmath>
#include <memory>

#include "runtime.cuh"   // (pulls comm.cuh in)
#include "runtime.cuh"
#include "visual.cuh"

struct lvba_visual_problem {
  int M = 0;
  long long T = 0, Tv = 0, nnz = 0, n_pairs = 0;
  int n_rows = 0, n_batches = 0, device = 0, fixed_cam = 0;
  std::vector<int> cam_of_row;
  cudaStream_t stream = nullptr;
  double intr[8];
  double sigma_px = 1, sigma_pl = 1;
  lvba::DevBuf<int> trk_ptr, trk_id, batch_trk, obs_cam, obs_row, d_cam_of_row;
  lvba::DevBuf<long long> batch_pair;
  lvba::DevBuf<unsigned> pairs;
  lvba::DevBuf<float2> obs_uv;
  lvba::DevBuf<double> plane;
  lvba::DevBuf<double> q, t, X, qc, tc, Xc;         // state and candidate
  lvba::DevBuf<double> q0, t0, X0;                  // state given at create (lvba_visual_reset_state)
  lvba::DevBuf<double> S, rhs, y, dadd, cam_colsq, cam_grad, s_cam, s_pt, pt_colsq;
  lvba::DevBuf<double> batch_cost, batch_gmax, batch_out, cam_out, scal, cam_step, pt_step;
  lvba::Envelope env;
  lvba::EnvSolver solver;
  lvba::PhaseTimers timers;
  double* h_scal = nullptr;
  int64_t launches = 0, h2d = 0, d2h = 0;
  double ms_setup = 0.0;
  // LM state
  lvba_visual_opts opts;
  double radius = 1e4, nu = 2.0, cost = 0.0, cost_first = 0.0;
  bool have_scale = false, have_first = false, converged = false;
  int iters = 0, accepted = 0, builds = 0, invalid = 0, termination = LVBA_TERM_MAX_ITER;

  lvba::VisualView view() const {
    lvba::VisualView v;
    v.n_batches = n_batches; v.trk_ptr = trk_ptr.p; v.trk_id = trk_id.p; v.batch_trk = batch_trk.p;
    v.batch_pair = batch_pair.p; v.pairs = pairs.p; v.obs_cam = obs_cam.p; v.obs_row = obs_row.p;
    v.obs_uv = obs_uv.p; v.plane = plane.p;
    for (int i = 0; i < 8; ++i) v.intr[i] = intr[i];
    v.inv_sigma_px = 1.0 / sigma_px;
    v.inv_sigma_pl = 1.0 / std::max(1e-9, sigma_pl);        // utils.hpp:131
    return v;
  }
  lvba::VisualState state() const { return lvba::VisualState{q.p, t.p, X.p}; }
  lvba::VisualState cand() const { return lvba::VisualState{qc.p, tc.p, Xc.p}; }
  ~lvba_visual_problem() {
    if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }     // buffers (members) must be idle when parked
    lvba::pinned_pool().give(h_scal, 16 * sizeof(double));                         // after the drain
  }
};

namespace lvba {

// ---- set-up kernels: the per-observation arrays and the pair table are derived on the device from the caller's arrays
// (uploaded as they are: one DMA each from the caller's — ideally pinned — memory) instead of being gathered on the host
// and pushed through pageable staging copies (config C: 16 MB of observations + 18 MB of pair words).
// One thread per local landmark k (source landmark trk_id[k]): its observations, camera rows and plane.
__global__ void visual_gather_kernel(long long Tv, const int* __restrict__ trk_id, const int* __restrict__ trk_ptr,
                                     const long long* __restrict__ obs_ptr, const int* __restrict__ raw_cam,
                                     const float2* __restrict__ raw_uv, const double* __restrict__ raw_plane,
                                     const int* __restrict__ row_of_cam, int* __restrict__ l_cam, int* __restrict__ l_row,
                                     float2* __restrict__ l_uv, double* __restrict__ l_plane) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Tv) return;
  const long long i = trk_id[k];
#pragma unroll
  for (int j = 0; j < 4; ++j) l_plane[4 * k + j] = raw_plane[4 * i + j];
  long long w = trk_ptr[k];
  for (long long q = obs_ptr[i]; q < obs_ptr[i + 1]; ++q, ++w) {
    const int c = raw_cam[q];
    l_cam[w] = c; l_row[w] = row_of_cam[c]; l_uv[w] = raw_uv[q];
  }
}

// The camera-pair words of one batch in the order visual_build_kernel walks them (landmarks ascending, x < y inside a
// landmark; hi = the larger reduced row; a camera observed twice contributes both orderings).  Thread = landmark of the batch;
// count == true writes the number of words of the batch to batch_pair[b + 1], otherwise the words go to pairs + batch_pair[b].
template <bool kCount>
__global__ void __launch_bounds__(kSlots)
visual_pairs_kernel(const int* __restrict__ trk_ptr, const int* __restrict__ batch_trk, const int* __restrict__ obs_row,
                    long long* __restrict__ batch_pair, unsigned* __restrict__ pairs) {
  __shared__ long long off[kMaxTrkPerBatch + 1];
  const int b = blockIdx.x, lt = threadIdx.x;
  const int k0 = batch_trk[b], nt = batch_trk[b + 1] - k0, sbase = trk_ptr[k0];
  int lo = 0, hi = 0;
  long long cnt = 0;
  if (lt < nt) {
    lo = trk_ptr[k0 + lt] - sbase; hi = trk_ptr[k0 + lt + 1] - sbase;
    for (int x = lo; x < hi; ++x) {
      const int rx = obs_row[sbase + x];
      if (rx < 0) continue;
      for (int y = x + 1; y < hi; ++y) {
        const int ry = obs_row[sbase + y];
        if (ry < 0) continue;
        cnt += (rx == ry) ? 2 : 1;
      }
    }
    off[lt + 1] = cnt;
  }
  if (lt == 0) off[0] = 0;
  __syncthreads();
  if (lt == 0) for (int i = 0; i < nt; ++i) off[i + 1] += off[i];
  __syncthreads();
  if (kCount) {
    if (lt == 0) batch_pair[b + 1] = off[nt];
    return;
  }
  if (lt >= nt) return;
  unsigned* dst = pairs + batch_pair[b] + off[lt];
  const unsigned tl = (unsigned)lt << 16;
  for (int x = lo; x < hi; ++x) {
    const int rx = obs_row[sbase + x];
    if (rx < 0) continue;
    for (int y = x + 1; y < hi; ++y) {
      const int ry = obs_row[sbase + y];
      if (ry < 0) continue;
      if (rx > ry) *dst++ = (unsigned)x | ((unsigned)y << 8) | tl;
      else if (ry > rx) *dst++ = (unsigned)y | ((unsigned)x << 8) | tl;
      else { *dst++ = (unsigned)x | ((unsigned)y << 8) | tl; *dst++ = (unsigned)y | ((unsigned)x << 8) | tl; }
    }
  }
}

// in-place inclusive scan of batch_pair[1..n] by one block (n_batches is a few ten thousand)
__global__ void visual_pair_scan_kernel(int n, long long* __restrict__ batch_pair) {
  __shared__ long long part[1024];
  const int tid = threadIdx.x, per = (n + 1023) / 1024;
  const int i0 = 1 + tid * per, i1 = min(n + 1, i0 + per);
  long long s = 0;
  for (int i = i0; i < i1; ++i) s += batch_pair[i];
  part[tid] = s;
  __syncthreads();
  if (tid == 0) { long long run = 0; for (int w = 0; w < 1024; ++w) { const long long v = part[w]; part[w] = run; run += v; } batch_pair[0] = 0; }
  __syncthreads();
  long long run = part[tid];
  for (int i = i0; i < i1; ++i) { run += batch_pair[i]; batch_pair[i] = run; }
}

inline bool plane_valid(const double* p) {   // has_valid_plane, src/lvba_system.cpp:1598
  for (int i = 0; i < 4; ++i) if (!std::isfinite(p[i])) return false;
  return std::fabs(p[0]) > 1e-6 || std::fabs(p[1]) > 1e-6 || std::fabs(p[2]) > 1e-6;
}

inline int visual_create_impl(int32_t M, int64_t T, const double* q, const double* t, const double* X,
                              const double* plane_nd, const int64_t* obs_ptr, const int32_t* obs_cam,
                              const float* obs_uv, const double intr[8], double sigma_px, double sigma_plane,
                              int32_t fixed_cam, int32_t device, lvba_visual_problem** out) {
  if (!out) return fail(LVBA_ERR_INVALID_ARG, "out is null");
  *out = nullptr;
  if (M <= 0 || T < 0) return fail(LVBA_ERR_INVALID_ARG, "M=%d T=%lld", M, (long long)T);
  if (!q || !t || !intr || !obs_ptr || (T > 0 && (!X || !plane_nd || !obs_cam || !obs_uv))) return fail(LVBA_ERR_INVALID_ARG, "null input pointer");
  if (!(sigma_px > 0)) return fail(LVBA_ERR_INVALID_ARG, "sigma_px must be positive");
  if (obs_ptr[0] != 0) return fail(LVBA_ERR_INVALID_ARG, "obs_ptr[0] must be 0");
  {
    int64_t bad_i[kMaxSetupThreads], bad_s[kMaxSetupThreads];
    for (int w = 0; w < kMaxSetupThreads; ++w) { bad_i[w] = -1; bad_s[w] = -1; }
    parallel_chunks(T, 1 << 14, [&](int64_t i0, int64_t i1, int w) {
      for (int64_t i = i0; i < i1; ++i) {
        if (obs_ptr[i + 1] < obs_ptr[i]) { bad_i[w] = i; return; }
        for (int64_t s = obs_ptr[i]; s < obs_ptr[i + 1]; ++s)
          if (obs_cam[s] < 0 || obs_cam[s] >= M) { bad_i[w] = i; bad_s[w] = s; return; }
      }
    });
    for (int w = 0; w < kMaxSetupThreads; ++w) {
      if (bad_i[w] < 0) continue;
      if (bad_s[w] < 0) return fail(LVBA_ERR_INVALID_ARG, "obs_ptr not monotone at %lld", (long long)bad_i[w]);
      return fail(LVBA_ERR_INVALID_ARG, "obs_cam[%lld]=%d out of [0,%d)", (long long)bad_s[w], obs_cam[bad_s[w]], M);
    }
  }
  LVBA_TRY(select_device(device));
  const double t_begin = wall_ms();
  const bool tlog = getenv("LVBA_SETUP_TIMING") != nullptr;
  double tprev = t_begin;
  auto lap = [&](const char* what) { if (tlog) { cudaStreamSynchronize(nullptr); const double tn = wall_ms(); fprintf(stderr, "[visual setup] %-26s %8.2f ms\n", what, tn - tprev); tprev = tn; } };
  std::unique_ptr<lvba_visual_problem> P(new lvba_visual_problem());
  P->M = M; P->T = T; P->fixed_cam = fixed_cam;
  for (int i = 0; i < 8; ++i) P->intr[i] = intr[i];
  P->sigma_px = sigma_px; P->sigma_pl = sigma_plane;
  LVBA_CUDA(cudaGetDevice(&P->device));
  LVBA_CUDA(cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking));
  P->timers.stream = P->stream;
  LVBA_TRY(pinned_pool().take(16 * sizeof(double), (void**)&P->h_scal));
  cudaStream_t s = P->stream;

  // ---- valid landmarks, active cameras (src/lvba_system.cpp:1582-1583, 1598-1603)
  std::vector<int64_t> valid;
  std::vector<char> cam_used(M, 0);
  {
    std::vector<char> is_valid((size_t)T, 0), used_w((size_t)kMaxSetupThreads * (size_t)M, 0);
    int64_t too_long[kMaxSetupThreads];
    for (int w = 0; w < kMaxSetupThreads; ++w) too_long[w] = -1;
    parallel_chunks(T, 1 << 13, [&](int64_t i0, int64_t i1, int w) {
      char* used = used_w.data() + (size_t)w * (size_t)M;
      for (int64_t i = i0; i < i1; ++i) {
        if (!plane_valid(plane_nd + 4 * i)) continue;
        if (obs_ptr[i + 1] - obs_ptr[i] > kSlots) { if (too_long[w] < 0) too_long[w] = i; continue; }
        is_valid[(size_t)i] = 1;
        for (int64_t q_ = obs_ptr[i]; q_ < obs_ptr[i + 1]; ++q_) used[obs_cam[q_]] = 1;
      }
    });
    for (int w = 0; w < kMaxSetupThreads; ++w)
      if (too_long[w] >= 0) {
        const int64_t i = too_long[w];
        return fail(LVBA_ERR_UNSUPPORTED, "landmark %lld has %lld observations; this build handles <= %d", (long long)i, (long long)(obs_ptr[i + 1] - obs_ptr[i]), kSlots);
      }
    for (int w = 0; w < kMaxSetupThreads; ++w)
      for (int c = 0; c < M; ++c) cam_used[c] |= used_w[(size_t)w * (size_t)M + c];
    for (int64_t i = 0; i < T; ++i) if (is_valid[(size_t)i]) valid.push_back(i);
  }
  if (fixed_cam >= 0 && fixed_cam < M) cam_used[fixed_cam] = 0;
  std::vector<int> row_of_cam(M, -1);
  for (int c = 0; c < M; ++c) if (cam_used[c]) { row_of_cam[c] = (int)P->cam_of_row.size(); P->cam_of_row.push_back(c); }
  P->n_rows = (int)P->cam_of_row.size();
  const int64_t Tv_all = (int64_t)valid.size();

  lap("valid landmarks / rows");
  // ---- envelope of the reduced camera system over ALL valid landmarks
  std::vector<int> first_raw(std::max(P->n_rows, 1));
  for (int r = 0; r < P->n_rows; ++r) first_raw[r] = r;
  std::vector<int> min_row(Tv_all, 0), min_sys_row(Tv_all, 0);
  {
    const size_t nr = first_raw.size();
    std::vector<int> first_w((size_t)kMaxSetupThreads * nr);
    for (int w = 0; w < kMaxSetupThreads; ++w) std::copy(first_raw.begin(), first_raw.end(), first_w.begin() + (size_t)w * nr);
    parallel_chunks(Tv_all, 1 << 13, [&](int64_t k0, int64_t k1, int w) {
      int* fr = first_w.data() + (size_t)w * nr;
      for (int64_t k = k0; k < k1; ++k) {
        const int64_t i = valid[k];
        int m = INT32_MAX, mcam = INT32_MAX;
        for (int64_t q_ = obs_ptr[i]; q_ < obs_ptr[i + 1]; ++q_) {
          const int r = row_of_cam[obs_cam[q_]];
          if (r >= 0) m = std::min(m, r);
          mcam = std::min(mcam, (int)obs_cam[q_]);
        }
        min_row[k] = (mcam == INT32_MAX) ? 0 : mcam;      // shard key: lowest camera index
        min_sys_row[k] = (m == INT32_MAX) ? 0 : m;        // row-owned reduced system: lowest row of the landmark's clique
        if (m == INT32_MAX) continue;
        for (int64_t q_ = obs_ptr[i]; q_ < obs_ptr[i + 1]; ++q_) {
          const int r = row_of_cam[obs_cam[q_]];
          if (r >= 0) fr[r] = std::min(fr[r], m);
        }
      }
    });
    for (int w = 0; w < kMaxSetupThreads; ++w)
      for (size_t r = 0; r < nr; ++r) first_raw[r] = std::min(first_raw[r], first_w[(size_t)w * nr + r]);
  }
  if (P->n_rows > 0) {
    first_raw.resize(P->n_rows);
    LVBA_TRY(P->env.build(first_raw, s, &P->h2d));
    LVBA_TRY(P->solver.prepare(P->env, s));
  }

  lap("envelope + solver prepare");
  // ---- shard: landmark -> owner of its lowest camera index
  Comm& cm = comm();
  std::vector<int64_t> mine;
  if (!cm.active()) mine.swap(valid);                          // one rank: every valid landmark
  else
    for (int64_t k = 0; k < Tv_all; ++k)
      if ((P->solver.dist() ? P->solver.dist_owner(min_sys_row[k]) : shard_owner(min_row[k], M, cm.n_ranks)) == cm.rank)
        mine.push_back(valid[k]);
  const int64_t Tv = (int64_t)mine.size();
  P->Tv = Tv;
  std::vector<int> trk_ptr(Tv + 1, 0), trk_id(Tv);
  for (int64_t k = 0; k < Tv; ++k) {
    trk_id[k] = (int)mine[k];
    trk_ptr[k + 1] = trk_ptr[k] + (int)(obs_ptr[mine[k] + 1] - obs_ptr[mine[k]]);
  }
  const long long nnz = trk_ptr[Tv];
  P->nnz = nnz;
  lap("landmark list");
  // ---- batches
  std::vector<int> batch_trk{0};
  {
    int ns = 0, nt = 0;
    for (int64_t k = 0; k < Tv; ++k) {
      const int L = trk_ptr[k + 1] - trk_ptr[k];
      if (nt > 0 && (ns + L > kSlots || nt + 1 > kMaxTrkPerBatch)) { batch_trk.push_back((int)k); ns = 0; nt = 0; }
      ns += L; ++nt;
    }
    if (Tv > 0) batch_trk.push_back((int)Tv);
  }
  P->n_batches = (int)batch_trk.size() - 1;
  lap("batches");

  // ---- upload the caller's arrays as they are; gather the per-observation arrays and build the pair table on the device
  LVBA_TRY(P->trk_ptr.upload(trk_ptr, s, &P->h2d));
  LVBA_TRY(P->batch_pair.alloc((size_t)P->n_batches + 1));
  LVBA_TRY(P->batch_pair.zero(s));
  if (Tv > 0) {
    const int64_t nnz_all = obs_ptr[T];
    DevBuf<long long> d_obs_ptr;
    DevBuf<int> d_raw_cam, d_row_of_cam;
    DevBuf<float2> d_raw_uv;
    DevBuf<double> d_raw_plane;
    static_assert(sizeof(long long) == sizeof(int64_t), "obs_ptr is uploaded as it is");
    LVBA_TRY(d_obs_ptr.upload(reinterpret_cast<const long long*>(obs_ptr), (size_t)T + 1, s, &P->h2d));
    LVBA_TRY(d_raw_cam.upload(obs_cam, (size_t)nnz_all, s, &P->h2d));
    LVBA_TRY(d_raw_uv.upload(reinterpret_cast<const float2*>(obs_uv), (size_t)nnz_all, s, &P->h2d));
    LVBA_TRY(d_raw_plane.upload(plane_nd, (size_t)T * 4, s, &P->h2d));
    LVBA_TRY(d_row_of_cam.upload(row_of_cam, s, &P->h2d));
    LVBA_TRY(P->trk_id.upload(trk_id, s, &P->h2d));
    LVBA_TRY(P->obs_cam.alloc((size_t)nnz)); LVBA_TRY(P->obs_row.alloc((size_t)nnz)); LVBA_TRY(P->obs_uv.alloc((size_t)nnz));
    LVBA_TRY(P->plane.alloc((size_t)Tv * 4));
    LVBA_TRY(P->batch_trk.upload(batch_trk, s, &P->h2d));
    visual_gather_kernel<<<(unsigned)((Tv + 127) / 128), 128, 0, s>>>(Tv, P->trk_id.p, P->trk_ptr.p, d_obs_ptr.p, d_raw_cam.p, d_raw_uv.p,
                                                                     d_raw_plane.p, d_row_of_cam.p, P->obs_cam.p, P->obs_row.p, P->obs_uv.p, P->plane.p);
    visual_pairs_kernel<true><<<P->n_batches, kSlots, 0, s>>>(P->trk_ptr.p, P->batch_trk.p, P->obs_row.p, P->batch_pair.p, nullptr);
    visual_pair_scan_kernel<<<1, 1024, 0, s>>>(P->n_batches, P->batch_pair.p);
    long long np = 0;
    LVBA_CUDA(cudaMemcpyAsync(&np, P->batch_pair.p + P->n_batches, sizeof(long long), cudaMemcpyDeviceToHost, s));
    LVBA_CUDA(cudaStreamSynchronize(s));          // the temporaries above may now go back to the pool
    P->n_pairs = np;
    if (np > 0) {
      LVBA_TRY(P->pairs.alloc((size_t)np));
      visual_pairs_kernel<false><<<P->n_batches, kSlots, 0, s>>>(P->trk_ptr.p, P->batch_trk.p, P->obs_row.p, P->batch_pair.p, P->pairs.p);
    }
    LVBA_CUDA(cudaGetLastError());
    P->launches += 4;
  } else {
    LVBA_TRY(P->batch_trk.upload(batch_trk, s, &P->h2d));
    P->n_pairs = 0;
  }
  lap("device gather + pair table");
  if (P->n_rows > 0) LVBA_TRY(P->d_cam_of_row.upload(P->cam_of_row, s, &P->h2d));
  LVBA_TRY(P->q.upload(q, (size_t)M * 4, s, &P->h2d));
  LVBA_TRY(P->t.upload(t, (size_t)M * 3, s, &P->h2d));
  LVBA_TRY(P->X.upload(X, (size_t)T * 3, s, &P->h2d));
  {   // restore point and candidate start as device-side copies of what just went up (no second and third trip over PCIe)
    auto dup = [&](DevBuf<double>& dst, const DevBuf<double>& src, size_t count) -> int {
      LVBA_TRY(dst.alloc(count));
      if (count) LVBA_CUDA(cudaMemcpyAsync(dst.p, src.p, count * sizeof(double), cudaMemcpyDeviceToDevice, s));
      return LVBA_OK;
    };
    LVBA_TRY(dup(P->q0, P->q, (size_t)M * 4)); LVBA_TRY(dup(P->t0, P->t, (size_t)M * 3)); LVBA_TRY(dup(P->X0, P->X, (size_t)T * 3));
    LVBA_TRY(dup(P->qc, P->q, (size_t)M * 4)); LVBA_TRY(dup(P->tc, P->t, (size_t)M * 3)); LVBA_TRY(dup(P->Xc, P->X, (size_t)T * 3));
  }
  const size_t n6 = (size_t)std::max(P->n_rows, 1) * 6;
  LVBA_TRY(P->S.alloc((size_t)std::max<long long>(P->env.nblocks, 1) * 36));
  LVBA_TRY(P->rhs.alloc(n6)); LVBA_TRY(P->y.alloc(n6)); LVBA_TRY(P->dadd.alloc(n6));
  LVBA_TRY(P->cam_colsq.alloc(n6)); LVBA_TRY(P->cam_grad.alloc(n6)); LVBA_TRY(P->s_cam.alloc(n6));
  LVBA_TRY(P->s_pt.alloc((size_t)std::max<int64_t>(Tv, 1) * 3));
  LVBA_TRY(P->pt_colsq.alloc((size_t)std::max<int64_t>(Tv, 1) * 3));
  const size_t nb = (size_t)std::max(P->n_batches, 1);
  LVBA_TRY(P->batch_cost.alloc(nb)); LVBA_TRY(P->batch_gmax.alloc(nb)); LVBA_TRY(P->batch_out.alloc(nb * 4));
  LVBA_TRY(P->cam_out.alloc((size_t)((P->n_rows + 127) / 128 + 1) * 2));
  LVBA_TRY(P->scal.alloc(16)); LVBA_TRY(P->scal.zero(s));
  LVBA_TRY(P->cam_step.alloc((size_t)M * 6)); LVBA_TRY(P->pt_step.alloc((size_t)std::max<int64_t>(T, 1) * 3));
  LVBA_CUDA(cudaFuncSetAttribute(visual_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)visual_build_smem_bytes()));
  LVBA_CUDA(cudaFuncSetAttribute(visual_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)visual_backsub_smem_bytes()));
  LVBA_CUDA(cudaStreamSynchronize(s));
  lap("uploads + allocations");
  lvba_visual_default_opts(&P->opts);
  P->ms_setup = wall_ms() - t_begin;
  *out = P.release();
  return LVBA_OK;
}

// Jacobi scaling vectors from the Jacobian at the current state (Ceres: once, at iteration 0)
inline int visual_compute_scale(lvba_visual_problem* P, int enabled) {
  cudaStream_t s = P->stream;
  LVBA_TRY(P->cam_colsq.zero(s));
  if (P->n_batches > 0) {
    visual_colnorm_kernel<<<P->n_batches, kSlots, 0, s>>>(P->view(), P->state(), P->cam_colsq.p, P->pt_colsq.p);
    ++P->launches;
  }
  Comm& cm = comm();
  if (cm.active()) LVBA_TRY(cm.allreduce_sum(P->cam_colsq.p, (size_t)P->n_rows * 6, s));
  const long long nc = (long long)P->n_rows * 6, np = (long long)P->Tv * 3;
  if (nc > 0) { visual_scale_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, s>>>(nc, P->cam_colsq.p, enabled, P->s_cam.p); ++P->launches; }
  if (np > 0) { visual_scale_kernel<<<(unsigned)((np + 255) / 256), 256, 0, s>>>(np, P->pt_colsq.p, enabled, P->s_pt.p); ++P->launches; }
  LVBA_CUDA(cudaGetLastError());
  P->have_scale = true;
  return LVBA_OK;
}

inline VisualLM visual_lm_params(const lvba_visual_problem* P, double radius) {
  return VisualLM{radius, P->opts.min_lm_diagonal, P->opts.max_lm_diagonal, P->s_cam.p, P->s_pt.p};
}

// scal layout: [0] cost  [1] gmax  [2] model  [3] step^2 (pts)  [4] x^2 (pts)  [5] -  [6] step^2 (cams) [7] x^2 (cams)
//              [8] candidate cost
inline int visual_linearize_solve(lvba_visual_problem* P, double radius, bool want_steps) {
  cudaStream_t s = P->stream;
  const VisualLM lm = visual_lm_params(P, radius);
  const EnvView ev = P->env.view();
  Comm& cm = comm();
  P->timers.begin(PH_BUILD);
  LVBA_TRY(P->S.zero(s)); LVBA_TRY(P->rhs.zero(s)); LVBA_TRY(P->cam_colsq.zero(s)); LVBA_TRY(P->cam_grad.zero(s));
  LVBA_CUDA(cudaMemsetAsync(P->scal.p + 1, 0, sizeof(double), s));
  if (P->n_batches > 0) {
    visual_build_kernel<<<P->n_batches, kSlots, visual_build_smem_bytes(), s>>>(
        P->view(), ev, P->state(), lm, P->S.p, P->rhs.p, P->cam_colsq.p, P->cam_grad.p, P->batch_cost.p, P->batch_gmax.p);
    ++P->launches;
  }
  reduce_partials_kernel<<<1, 256, 0, s>>>(P->batch_cost.p, P->n_batches, P->scal.p + 0);
  reduce_max_kernel<<<1, 256, 0, s>>>(P->batch_gmax.p, P->n_batches, P->scal.p + 1);
  P->launches += 2;
  if (cm.active()) {
    // row-owned reduced camera system: see lidar_build_dev
    if (P->solver.dist()) LVBA_TRY(P->solver.exchange_rows(P->env, P->S.p, s, &P->launches));
    else LVBA_TRY(cm.allreduce_sum(P->S.p, (size_t)P->env.nblocks * 36, s));
    LVBA_TRY(cm.allreduce_sum(P->rhs.p, (size_t)P->n_rows * 6, s));
    LVBA_TRY(cm.allreduce_sum(P->cam_colsq.p, (size_t)P->n_rows * 6, s));
    LVBA_TRY(cm.allreduce_sum(P->cam_grad.p, (size_t)P->n_rows * 6, s));
    LVBA_TRY(cm.allreduce_sum(P->scal.p + 0, 1, s));
  }
  P->timers.end();
  ++P->builds;
  P->timers.begin(PH_SOLVE);
  if (P->n_rows > 0) {
    visual_cam_diag_kernel<<<1, 256, 0, s>>>(6 * P->n_rows, P->cam_colsq.p, P->cam_grad.p, P->s_cam.p, P->opts.min_lm_diagonal,
                                             P->opts.max_lm_diagonal, radius, P->dadd.p, P->scal.p + 5);
    ++P->launches;
    LVBA_CUDA(cudaMemcpyAsync(P->solver.z.p, P->rhs.p, (size_t)P->n_rows * 6 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    LVBA_TRY(P->solver.solve(P->env, P->S.p, P->dadd.p, P->y.p, s, &P->launches));
  }
  P->timers.end();
  P->timers.begin(PH_RESID);
  if (P->n_batches > 0) {
    visual_backsub_kernel<<<P->n_batches, kSlots, visual_backsub_smem_bytes(), s>>>(
        P->view(), P->state(), lm, P->y.p, P->Xc.p, want_steps ? P->pt_step.p : nullptr, P->batch_out.p);
    ++P->launches;
  }
  reduce_cols_kernel<<<1, 256, 0, s>>>(P->batch_out.p, P->n_batches, 4, 3, P->scal.p + 2);
  ++P->launches;
  const int ncb = (P->n_rows + 127) / 128;
  if (ncb > 0) {
    visual_cam_update_kernel<<<ncb, 128, 0, s>>>(P->n_rows, P->d_cam_of_row.p, P->q.p, P->t.p, P->y.p, P->s_cam.p, P->qc.p, P->tc.p,
                                                 want_steps ? P->cam_step.p : nullptr, P->cam_out.p);
    ++P->launches;
  }
  reduce_cols_kernel<<<1, 256, 0, s>>>(P->cam_out.p, ncb, 2, 2, P->scal.p + 6);
  ++P->launches;
  if (cm.active()) LVBA_TRY(cm.allreduce_sum(P->scal.p + 2, 3, s));   // model, step^2, x^2 of the landmark shard
  // candidate cost
  if (P->n_batches > 0) {
    visual_cost_kernel<<<P->n_batches, kSlots, 0, s>>>(P->view(), P->cand(), P->batch_cost.p);
    ++P->launches;
  }
  reduce_partials_kernel<<<1, 256, 0, s>>>(P->batch_cost.p, P->n_batches, P->scal.p + 8);
  ++P->launches;
  if (cm.active()) LVBA_TRY(cm.allreduce_sum(P->scal.p + 8, 1, s));
  P->timers.end();
  LVBA_CUDA(cudaGetLastError());
  LVBA_CUDA(cudaMemcpyAsync(P->h_scal, P->scal.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, s));
  LVBA_CUDA(cudaMemcpyAsync(P->h_scal + 15, P->solver.status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  LVBA_CUDA(cudaStreamSynchronize(s));
  P->d2h += 16 * sizeof(double);
  return LVBA_OK;
}

inline int visual_iterate_impl(lvba_visual_problem* P, int n_iter, lvba_summary* sum) {
  const double t0 = wall_ms();
  const int64_t l0 = P->launches, h0 = P->h2d, d0 = P->d2h;
  const int iters0 = P->iters, acc0 = P->accepted, builds0 = P->builds;
  const lvba_visual_opts& o = P->opts;
  if (!P->have_scale) LVBA_TRY(visual_compute_scale(P, o.jacobi_scaling));
  for (int it = 0; it < n_iter && !P->converged; ++it) {
    LVBA_TRY(visual_linearize_solve(P, P->radius, false));
    const double* h = P->h_scal;
    P->cost = h[0];
    if (!P->have_first) { P->cost_first = P->cost; P->have_first = true; }
    const double gmax = std::max(h[1], h[5]);
    if (o.gradient_tolerance >= 0 && gmax <= o.gradient_tolerance) { P->converged = true; P->termination = LVBA_TERM_GRADIENT_TOL; break; }
    ++P->iters;
    const double model = h[2];
    const double cand = h[8];
    const int fstat = *reinterpret_cast<const int*>(h + 15);
    const double step_norm = std::sqrt(h[3] + h[6]), x_norm = std::sqrt(h[4] + h[7]);
    const bool valid = fstat == 0 && std::isfinite(model) && std::isfinite(step_norm) && model > 0.0;
    if (o.verbose)
      fprintf(stderr, "[lvba visual] iter %d: cost %.9g cand %.9g model %.6g radius %.3g step %.3g valid %d\n", P->iters, P->cost, cand, model, P->radius, step_norm, (int)valid);
    if (!valid) {                                    // LevenbergMarquardtStrategy::StepIsInvalid
      ++P->invalid;
      P->radius *= 0.5;
      if (P->invalid >= 5) { P->converged = true; P->termination = LVBA_TERM_INVALID_STEPS; }
      continue;
    }
    P->invalid = 0;
    const double rho = (P->cost - cand) / model;
    if (o.parameter_tolerance >= 0 && step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) {
      P->converged = true; P->termination = LVBA_TERM_PARAMETER_TOL; break;
    }
    if (o.function_tolerance >= 0 && std::fabs(P->cost - cand) <= o.function_tolerance * P->cost) {
      P->converged = true; P->termination = LVBA_TERM_FUNCTION_TOL; break;     // Ceres returns x, not the candidate
    }
    if (std::isfinite(cand) && rho > o.min_relative_decrease) {             // StepAccepted
      std::swap(P->q.p, P->qc.p); std::swap(P->t.p, P->tc.p); std::swap(P->X.p, P->Xc.p);
      // keep untouched entries (constant cameras, skipped landmarks) identical in both buffers: they never change
      P->cost = cand;
      ++P->accepted;
      P->radius = std::min(o.max_radius, P->radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
      P->nu = 2.0;
    } else {                                                                  // StepRejected
      P->radius /= P->nu;
      P->nu *= 2.0;
      if (P->radius < o.min_radius) { P->converged = true; P->termination = LVBA_TERM_RADIUS; }
    }
  }
  if (sum) {
    memset(sum, 0, sizeof *sum);
    LVBA_CUDA(cudaStreamSynchronize(P->stream));
    double ms[PH_COUNT] = {0, 0, 0};
    P->timers.collect(ms);
    sum->iterations = P->iters - iters0; sum->accepted = P->accepted - acc0; sum->hessian_builds = P->builds - builds0;
    sum->termination = P->termination; sum->cost_first = P->cost_first; sum->cost_last = P->cost;
    sum->damping_last = P->radius; sum->ms_total = wall_ms() - t0;
    sum->ms_build = ms[PH_BUILD]; sum->ms_solve = ms[PH_SOLVE]; sum->ms_residual = ms[PH_RESID];
    sum->kernel_launches = P->launches - l0; sum->h2d_bytes = P->h2d - h0; sum->d2h_bytes = P->d2h - d0;
  }
  return LVBA_OK;
}

}  // namespace lvba

extern "C" {

void lvba_visual_default_opts(lvba_visual_opts* o) {
  if (!o) return;
  o->max_iter = 50; o->initial_radius = 1e4; o->max_radius = 1e16; o->min_radius = 1e-32;
  o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32; o->min_relative_decrease = 1e-3;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->jacobi_scaling = 1; o->device = -1; o->verbose = 0;
}

int lvba_visual_create(int32_t M, int64_t T, const double* q, const double* t, const double* X, const double* plane_nd,
                       const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, const double intr[8],
                       double sigma_px, double sigma_plane, int32_t fixed_cam, int32_t device, lvba_visual_problem** out) {
  try { return lvba::visual_create_impl(M, T, q, t, X, plane_nd, obs_ptr, obs_cam, obs_uv, intr, sigma_px, sigma_plane, fixed_cam, device, out); }
  catch (const std::bad_alloc&) { return lvba::fail(LVBA_ERR_NOMEM, "host allocation failed"); }
  catch (...) { return lvba::fail(LVBA_ERR_INVALID_ARG, "unexpected exception in lvba_visual_create"); }
}

int lvba_visual_destroy(lvba_visual_problem* p) LVBA_ABI_BEGIN {
  if (!p) return LVBA_OK;
  cudaSetDevice(p->device);
  delete p;
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_destroy")

int lvba_visual_set_state(lvba_visual_problem* p, const double* q, const double* t, const double* X) LVBA_ABI_BEGIN {
  if (!p || !q || !t || (p->T > 0 && !X)) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  LVBA_CUDA(cudaSetDevice(p->device));
  LVBA_TRY(p->q.upload(q, (size_t)p->M * 4, p->stream, &p->h2d)); LVBA_TRY(p->qc.upload(q, (size_t)p->M * 4, p->stream));
  LVBA_TRY(p->t.upload(t, (size_t)p->M * 3, p->stream, &p->h2d)); LVBA_TRY(p->tc.upload(t, (size_t)p->M * 3, p->stream));
  LVBA_TRY(p->X.upload(X, (size_t)p->T * 3, p->stream, &p->h2d)); LVBA_TRY(p->Xc.upload(X, (size_t)p->T * 3, p->stream));
  LVBA_CUDA(cudaStreamSynchronize(p->stream));
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_set_state")

int lvba_visual_get_state(lvba_visual_problem* p, double* q, double* t, double* X) LVBA_ABI_BEGIN {
  if (!p) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  LVBA_CUDA(cudaSetDevice(p->device));
  if (q) LVBA_CUDA(cudaMemcpyAsync(q, p->q.p, (size_t)p->M * 4 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (t) LVBA_CUDA(cudaMemcpyAsync(t, p->t.p, (size_t)p->M * 3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (X && p->T > 0) {
    if (lvba::comm().active()) {
      // landmarks are sharded: gather every rank's updates (disjoint) with one all-reduce of the deltas
      lvba::DevBuf<double> delta, merged;
      const long long n3 = (long long)p->T * 3;
      LVBA_TRY(delta.alloc((size_t)n3)); LVBA_TRY(merged.alloc((size_t)n3));
      LVBA_TRY(delta.zero(p->stream));
      if (p->Tv > 0) { lvba::visual_delta_kernel<<<(unsigned)((p->Tv * 3 + 255) / 256), 256, 0, p->stream>>>(p->Tv, p->trk_id.p, p->X.p, p->X0.p, delta.p); ++p->launches; }
      LVBA_TRY(lvba::comm().allreduce_sum(delta.p, (size_t)n3, p->stream));
      lvba::visual_add_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, p->stream>>>(n3, p->X0.p, delta.p, merged.p); ++p->launches;
      LVBA_CUDA(cudaMemcpyAsync(X, merged.p, (size_t)n3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
      LVBA_CUDA(cudaStreamSynchronize(p->stream));
    } else {
      LVBA_CUDA(cudaMemcpyAsync(X, p->X.p, (size_t)p->T * 3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    }
  }
  LVBA_CUDA(cudaStreamSynchronize(p->stream));
  p->d2h += (int64_t)p->M * 56 + p->T * 24;
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_get_state")

int lvba_visual_cost(lvba_visual_problem* p, double* cost) LVBA_ABI_BEGIN {
  if (!p || !cost) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  LVBA_CUDA(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  if (p->n_batches > 0) { lvba::visual_cost_kernel<<<p->n_batches, lvba::kSlots, 0, s>>>(p->view(), p->state(), p->batch_cost.p); ++p->launches; }
  lvba::reduce_partials_kernel<<<1, 256, 0, s>>>(p->batch_cost.p, p->n_batches, p->scal.p + 8);
  ++p->launches;
  if (lvba::comm().active()) LVBA_TRY(lvba::comm().allreduce_sum(p->scal.p + 8, 1, s));
  LVBA_CUDA(cudaMemcpyAsync(p->h_scal, p->scal.p + 8, sizeof(double), cudaMemcpyDeviceToHost, s));
  LVBA_CUDA(cudaStreamSynchronize(s));
  *cost = p->h_scal[0];
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_cost")

int lvba_visual_step(lvba_visual_problem* p, double radius, int32_t jacobi_scaling, int32_t recompute_scale,
                     double* cam_step, double* pt_step, double* model_cost_change, double* cost) LVBA_ABI_BEGIN {
  if (!p || !(radius > 0)) return lvba::fail(LVBA_ERR_INVALID_ARG, "bad argument");
  LVBA_CUDA(cudaSetDevice(p->device));
  if (recompute_scale || !p->have_scale) LVBA_TRY(lvba::visual_compute_scale(p, jacobi_scaling));
  LVBA_CUDA(cudaMemsetAsync(p->cam_step.p, 0, (size_t)p->M * 6 * sizeof(double), p->stream));
  if (p->T > 0) LVBA_CUDA(cudaMemsetAsync(p->pt_step.p, 0, (size_t)p->T * 3 * sizeof(double), p->stream));
  LVBA_TRY(lvba::visual_linearize_solve(p, radius, true));
  if (cam_step) LVBA_CUDA(cudaMemcpyAsync(cam_step, p->cam_step.p, (size_t)p->M * 6 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (pt_step && p->T > 0) LVBA_CUDA(cudaMemcpyAsync(pt_step, p->pt_step.p, (size_t)p->T * 3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  LVBA_CUDA(cudaStreamSynchronize(p->stream));
  if (model_cost_change) *model_cost_change = p->h_scal[2];
  if (cost) *cost = p->h_scal[0];
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_step")

int lvba_visual_structure(lvba_visual_problem* p, int32_t* n_active, int32_t* cam_of_row, int64_t* nblocks, int32_t* brow, int32_t* bcol) LVBA_ABI_BEGIN {
  if (!p) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  if (n_active) *n_active = p->n_rows;
  if (cam_of_row) for (int r = 0; r < p->n_rows; ++r) cam_of_row[r] = p->cam_of_row[r];
  if (nblocks) *nblocks = p->env.nblocks;
  if (brow && bcol)
    for (int r = 0; r < p->env.n; ++r)
      for (int c = p->env.first[r]; c <= r; ++c) {
        const long long b = p->env.row_start[r] + (c - p->env.first[r]);
        brow[b] = r; bcol[b] = c;
      }
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_structure")

int lvba_visual_get_system(lvba_visual_problem* p, double* rhs, double* blocks) LVBA_ABI_BEGIN {
  if (!p) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  LVBA_CUDA(cudaSetDevice(p->device));
  if (rhs && p->n_rows > 0) LVBA_CUDA(cudaMemcpyAsync(rhs, p->rhs.p, (size_t)p->n_rows * 6 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (blocks && p->env.nblocks > 0) LVBA_CUDA(cudaMemcpyAsync(blocks, p->S.p, (size_t)p->env.nblocks * 36 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  LVBA_CUDA(cudaStreamSynchronize(p->stream));
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_get_system")

int lvba_visual_reset_lm(lvba_visual_problem* p, const lvba_visual_opts* opts) LVBA_ABI_BEGIN {
  if (!p) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  if (opts) p->opts = *opts; else lvba_visual_default_opts(&p->opts);
  p->radius = p->opts.initial_radius; p->nu = 2.0; p->have_scale = false; p->have_first = false; p->converged = false;
  p->iters = p->accepted = p->builds = p->invalid = 0; p->termination = LVBA_TERM_MAX_ITER;
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_reset_lm")

int lvba_visual_reset_state(lvba_visual_problem* p) LVBA_ABI_BEGIN {
  if (!p) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  LVBA_CUDA(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  const size_t nq = (size_t)p->M * 4 * sizeof(double), nt = (size_t)p->M * 3 * sizeof(double), nx = (size_t)p->T * 3 * sizeof(double);
  LVBA_CUDA(cudaMemcpyAsync(p->q.p, p->q0.p, nq, cudaMemcpyDeviceToDevice, s)); LVBA_CUDA(cudaMemcpyAsync(p->qc.p, p->q0.p, nq, cudaMemcpyDeviceToDevice, s));
  LVBA_CUDA(cudaMemcpyAsync(p->t.p, p->t0.p, nt, cudaMemcpyDeviceToDevice, s)); LVBA_CUDA(cudaMemcpyAsync(p->tc.p, p->t0.p, nt, cudaMemcpyDeviceToDevice, s));
  if (nx) { LVBA_CUDA(cudaMemcpyAsync(p->X.p, p->X0.p, nx, cudaMemcpyDeviceToDevice, s)); LVBA_CUDA(cudaMemcpyAsync(p->Xc.p, p->X0.p, nx, cudaMemcpyDeviceToDevice, s)); }
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_reset_state")

int lvba_visual_iterate(lvba_visual_problem* p, int32_t n_iter, lvba_summary* summary) LVBA_ABI_BEGIN {
  if (!p || n_iter < 0) return lvba::fail(LVBA_ERR_INVALID_ARG, "bad argument");
  LVBA_CUDA(cudaSetDevice(p->device));
  return lvba::visual_iterate_impl(p, n_iter, summary);
} LVBA_ABI_END("lvba_visual_iterate")

int lvba_visual_counts(lvba_visual_problem* p, int64_t* nnz_valid, int64_t* n_valid_tracks, int64_t* n_blocks_env, int64_t* n_pairs) LVBA_ABI_BEGIN {
  if (!p) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  if (nnz_valid) *nnz_valid = p->nnz;
  if (n_valid_tracks) *n_valid_tracks = p->Tv;
  if (n_blocks_env) *n_blocks_env = p->env.nblocks;
  if (n_pairs) *n_pairs = p->n_pairs;
  return LVBA_OK;
} LVBA_ABI_END("lvba_visual_counts")

int lvba_visual_lm(int32_t M, int64_t T, double* q, double* t, double* X, const double* plane_nd, const int64_t* obs_ptr,
                   const int32_t* obs_cam, const float* obs_uv, const double intr[8], double sigma_px, double sigma_plane,
                   int32_t fixed_cam, const lvba_visual_opts* opts, lvba_summary* summary) LVBA_ABI_BEGIN {
  const double t0 = lvba::wall_ms();
  lvba_visual_opts o;
  if (opts) o = *opts; else lvba_visual_default_opts(&o);
  lvba_visual_problem* p = nullptr;
  int rc = lvba_visual_create(M, T, q, t, X, plane_nd, obs_ptr, obs_cam, obs_uv, intr, sigma_px, sigma_plane, fixed_cam, o.device, &p);
  if (rc != LVBA_OK) return rc;
  lvba_summary s;
  memset(&s, 0, sizeof s);
  rc = lvba_visual_reset_lm(p, &o);
  if (rc == LVBA_OK && p->Tv > 0) rc = lvba_visual_iterate(p, o.max_iter, &s);
  if (rc == LVBA_OK) rc = lvba_visual_get_state(p, q, t, X);   // write-back, src/lvba_system.cpp:1651-1665
  if (rc == LVBA_OK && summary) {
    *summary = s;
    summary->ms_setup = p->ms_setup; summary->kernel_launches = p->launches;
    summary->h2d_bytes = p->h2d; summary->d2h_bytes = p->d2h; summary->ms_total = lvba::wall_ms() - t0;
  }
  lvba_visual_destroy(p);
  return rc;
} LVBA_ABI_END("lvba_visual_lm")

}  // extern "C"
