// Bring-up harness for lhotse_b200/csrc/tc512.cuh (the tcgen05 two-stage DFT kernel).  Modes (one per process: a bad
// descriptor kills the context):
//   disc   layout discovery: one MMA with an index-valued operand against an identity operand in the K-major no-swizzle
//          layout round 1 verified (scripts/micro/umma_tf32_probe.cu); D then lists WHICH shared-memory word the tensor core
//          read as element (row, k) — checked against the formulas tc512.cuh assumes (MN-major 64 B swizzle A, K-major
//          128 B swizzle A and B, K-step advances)
//   rate   cycles per tcgen05.mma (M128 N{32,64,128,256} K8 tf32), one CTA and 2 CTAs on every SM
//   tc     the product kernel with DBG dumps on a small ragged batch: D1 / D2 / P of the first tile and every output
//          row against a float64 evaluation of the same tables
//   bench  2048 x 10 s cuts, CUDA events
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc512_probe tc512_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../lhotse_b200/csrc/tc512.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// ------------------------------------------------------------------------------------------------ discovery
// K-major no-swizzle canonical index (verified in round 1): core (row/8, k/4) at (row/8)*SBO + (k/4)*LBO, 8 rows x 16 B
__host__ __device__ inline int ns_index(int row, int k, int K) { return (row / 8) * (32 * (K / 4)) + (k / 4) * 32 + (row % 8) * 4 + (k % 4); }

struct DiscCfg {
  int which;            // 0: discover A, 1: discover B
  int pass;             // operand word i holds (i >> (11 * pass)) & 2047
  uint32_t lbo, sbo, layout, a_mn, b_mn, start_off;
  int words;            // operand words to fill
};

__global__ void __launch_bounds__(128) disc_kernel(DiscCfg c, float *D) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float *X = reinterpret_cast<float *>(sm);              // operand under test (up to 32 KB)
  float *I = reinterpret_cast<float *>(sm + 32768);      // identity operand, K-major no swizzle, 128 rows x K=8
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + 32768 + 4096);
  uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < c.words; i += 128) X[i] = (float)((i >> (11 * c.pass)) & 2047);
  for (int i = tid; i < 1024; i += 128) I[i] = 0.f;
  __syncthreads();
  if (tid < 8) I[ns_index(tid, tid, 8)] = 1.f;           // rows 0..7: delta(row, k)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(tc_smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot;
  if (tid == 0) {
    const uint64_t xd = tc_desc(tc_smem_u32(X) + c.start_off, c.lbo, c.sbo, c.layout);
    const uint64_t id = tc_desc(tc_smem_u32(I), 128, 256, 0);  // K = 8: two cores per 8-row group
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (c.a_mn << 15) | (c.b_mn << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    if (c.which == 0) tc_mma(tm, xd, id, idesc, 0);   // D[row][col] = sum_k X(row, k) I(col, k) = X(row, col) for col < 8
    else tc_mma(tm, id, xd, idesc, 0);                // D[row][col] = sum_k I(row, k) X(col, k) = X(col, k = row) for row < 8
    tc_commit(tc_smem_u32(bar));
  }
  tc_wait(tc_smem_u32(bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  tc_ld32(tm + ((uint32_t)(warp * 32) << 16), r);
  for (int n = 0; n < 32; ++n) D[tid * 32 + n] = __uint_as_float(r[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm));
}

static std::vector<int> run_disc(DiscCfg c) {  // returns the observed word index per D element [128][32]
  float *dD;
  CK(cudaMalloc(&dD, 128 * 32 * 4));
  std::vector<float> lo(128 * 32), hi(128 * 32);
  CK(cudaFuncSetAttribute(disc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
  for (int pass = 0; pass < 2; ++pass) {
    c.pass = pass;
    disc_kernel<<<1, 128, 40 * 1024>>>(c, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(pass ? hi.data() : lo.data(), dD, 128 * 32 * 4, cudaMemcpyDeviceToHost));
  }
  cudaFree(dD);
  std::vector<int> idx(128 * 32);
  for (int i = 0; i < 128 * 32; ++i) idx[i] = (int)lo[i] + ((int)hi[i] << 11);
  return idx;
}

static int mn32b_index(int row, int k) {  // tc512.cuh A1 (A1M = 1): row = 32 g + m, k = n1 within one K = 8 step (two 4-row atoms), floats
  const int g = row >> 5, m = row & 31, r = k & 3;
  return g * 1024 + (k >> 2) * 128 + r * 32 + ((((m >> 3) ^ r) & 3) << 3) + (m & 7);
}

static void dump_rows(const std::vector<int> &idx) {
  for (int row : {0, 1, 7, 8, 16, 31, 32, 64}) {
    printf("  row %3d:", row);
    for (int k = 0; k < 8; ++k) printf(" %5d", idx[row * 32 + k]);
    printf("\n");
  }
}

static int disc_main() {
  int bad_total = 0;
  {  // A, MN-major, SWIZZLE_128B_BASE32B (layout 1): LBO = 4096 (next 32 rows), SBO = 512 (next 4 k)
    struct Cand { uint32_t lbo, sbo, layout; const char *name; };
    const Cand cands[] = {{4096, 512, 1, "LBO=4096 SBO=512 layout=1"}, {512, 4096, 1, "LBO=512 SBO=4096 layout=1 (swapped)"}};
    for (const Cand &cd : cands)
      for (int ks = 0; ks < 2; ++ks) {
        DiscCfg c{0, 0, cd.lbo, cd.sbo, cd.layout, 1, 0, (uint32_t)(ks * 1024), 8192};
        auto idx = run_disc(c);
        int bad = 0;
        for (int row = 0; row < 128; ++row)
          for (int k = 0; k < 8; ++k) {
            const int exp = mn32b_index(row, k) + ks * 256;
            if (idx[row * 32 + k] != exp) ++bad;
          }
        printf("disc A MN-major BASE32B %s kstep %d: %s (%d mismatches)\n", cd.name, ks, bad ? "MISMATCH" : "OK", bad);
        if (bad) dump_rows(idx);
        if (cd.lbo == 4096) bad_total += bad;
      }
  }
  {  // A, K-major, 128-byte swizzle, SBO = 1024, K-step advance 32 B
    for (int ks = 0; ks < 4; ++ks) {
      DiscCfg c{0, 0, 16, 1024, 2, 0, 0, (uint32_t)(ks * 32), 4096};
      auto idx = run_disc(c);
      int bad = 0;
      for (int row = 0; row < 128; ++row)
        for (int k = 0; k < 8; ++k) {
          const int exp = tc_k128_index(row, 8 * ks + k);
          if (idx[row * 32 + k] != exp) { if (bad < 6) printf("  A k128 ks%d (row %d, k %d): observed word %d, expected %d\n", ks, row, k, idx[row * 32 + k], exp); ++bad; }
        }
      printf("disc A K-major SW128 kstep %d: %s (%d mismatches)\n", ks, bad ? "MISMATCH" : "OK", bad);
      bad_total += bad;
    }
  }
  {  // B, K-major, 128-byte swizzle (N = 32 rows)
    for (int ks = 0; ks < 4; ++ks) {
      DiscCfg c{1, 0, 16, 1024, 2, 0, 0, (uint32_t)(ks * 32), 1024};
      auto idx = run_disc(c);
      int bad = 0;
      for (int k = 0; k < 8; ++k)
        for (int col = 0; col < 32; ++col) {
          const int exp = tc_k128_index(col, 8 * ks + k);
          if (idx[k * 32 + col] != exp) { if (bad < 6) printf("  B k128 ks%d (col %d, k %d): observed word %d, expected %d\n", ks, col, k, idx[k * 32 + col], exp); ++bad; }
        }
      printf("disc B K-major SW128 kstep %d: %s (%d mismatches)\n", ks, bad ? "MISMATCH" : "OK", bad);
      bad_total += bad;
    }
  }
  printf("DISC %s\n", bad_total ? "MISMATCH" : "ALL OK");
  return bad_total ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ rate
// mode 0: A from shared memory, K-major 128-byte swizzle (what tc512 uses)   1: K-major 32-byte swizzle, 4 KB contiguous per K-step
// mode 2: K-major no swizzle, contiguous                                     3: A from TMEM (tcgen05.mma [d], [a], b-desc)
template <int NN, int MM>
__global__ void __launch_bounds__(128) rate_kernel(int nmma, int mode, long long *cycles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + 65536);
  uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += 128) reinterpret_cast<float *>(sm)[i] = 0.f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tc_smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot;
  long long t0 = 0;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);
    const uint32_t a = tc_smem_u32(sm), bb = tc_smem_u32(sm) + 32768;
    const uint64_t bd = tc_desc(bb, 16, 1024, 2);
    t0 = clock64();
    if (mode == 3) {
      for (int i = 0; i < nmma; ++i) {
        const uint32_t acc = i > 1, d = tm + ((2 * NN <= 224) ? (i & 1) * NN : 0), at = tm + 224 + (i & 3) * 8;
        const uint64_t b2 = bd + (uint64_t)((i & 3) * 2);
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
                     "  tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }\n"
                     ::"r"(d), "r"(at), "l"(b2), "r"(idesc), "r"(acc) : "memory");
      }
    } else {
      for (int i = 0; i < nmma; ++i) {
        uint64_t ad;
        if (mode == 0) ad = tc_desc(a + (i & 3) * 32, 16, 1024, 2);
        else if (mode == 1) ad = tc_desc(a + (i & 3) * 4096, 16, 256, 6);
        else ad = tc_desc(a + (i & 3) * 4096, 128, 256, 0);
        tc_mma(tm + ((2 * NN <= 224) ? (i & 1) * NN : 0), ad, bd + (uint64_t)((i & 3) * 2), idesc, i > 1);
      }
    }
    tc_commit(tc_smem_u32(bar));
  }
  tc_wait(tc_smem_u32(bar), 0);
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm));
}

template <int NN, int MM>
static void rate_one(int grid, int mode) {
  const int nmma = 2048;
  long long *d;
  CK(cudaMalloc(&d, grid * 8));
  CK(cudaFuncSetAttribute(rate_kernel<NN, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024));
  rate_kernel<NN, MM><<<grid, 128, 66 * 1024>>>(nmma, mode, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("rate M%d N%d mode %d: %s\n", MM, NN, mode, cudaGetErrorString(e)); exit(3); }
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0; double avg = 0;
  for (auto v : h) { mx = v > mx ? v : mx; avg += (double)v / grid; }
  const char *names[] = {"smem K-major SW128", "smem K-major SW32 contiguous", "smem K-major no-swizzle contiguous", "A in TMEM"};
  printf("rate M%-3d N%-3d K8 tf32 [%s], grid %3d: %.1f cycles/MMA avg (max %.1f) -> %.0f MAC/cycle/CTA\n", MM, NN, names[mode], grid, avg / nmma,
         (double)mx / nmma, (double)MM * NN * 8 / (avg / nmma));
  cudaFree(d);
}

static int rate_main(int mode) {
  for (int grid : {1, 296}) {
    rate_one<32, 128>(grid, mode); rate_one<64, 128>(grid, mode); rate_one<96, 128>(grid, mode); rate_one<128, 128>(grid, mode);
    rate_one<32, 64>(grid, mode); rate_one<64, 64>(grid, mode);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ full kernel
struct Problem {
  DevPlan p;
  std::vector<float> window, bank;  // bank [K][M]
  std::vector<int64_t> ns;
  std::vector<float> x;             // packed, 4-aligned starts
  std::vector<int64_t> meta;        // soff | n | row prefix | tile prefix | tile -> cut (int32)
  int64_t rows = 0, tiles = 0, span = 0;
};

static void make_problem(Problem &P, const std::vector<int64_t> &ns, int M, unsigned seed) {
  DevPlan &p = P.p;
  memset(&p, 0, sizeof(p));
  p.feature = B200FEAT_FBANK; p.L = 400; p.S = 160; p.N = 512; p.K = 257; p.M = M; p.C = 0; p.F = M;
  p.Nc = 256; p.packed = 1; p.pad_left = 120; p.pad_mode = 0; p.snip_edges = 0; p.remove_dc = 1;
  p.preemph = 0.97f; p.mel_floor = 1.1920929e-07f;
  P.window.resize(400);
  for (int i = 0; i < 400; ++i) P.window[i] = (float)pow(0.5 - 0.5 * cos(2.0 * M_PI * i / 399.0), 0.85);
  P.bank.assign((size_t)257 * M, 0.f);
  auto mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
  const double lo = mel(20.0), hi = mel(7600.0), d = (hi - lo) / (M + 1);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 256; ++k) {
      const double mk = mel(k * 16000.0 / 512.0), l = lo + m * d, c = l + d, r = l + 2 * d;
      const double w = fmin((mk - l) / (c - l), (r - mk) / (r - c));
      if (w > 0) P.bank[(size_t)k * M + m] = (float)w;
    }
  P.ns = ns;
  const int B = (int)ns.size();
  P.meta.assign(4 * B + 2, 0);
  int64_t cur = 0, rows = 0, tiles = 0;
  for (int i = 0; i < B; ++i) {
    cur = (cur + 3) / 4 * 4;
    P.meta[i] = cur; P.meta[B + i] = ns[i]; P.meta[2 * B + i] = rows; P.meta[3 * B + 1 + i] = tiles;
    cur += ns[i];
    const int64_t T = (ns[i] + 80) / 160;
    rows += T; tiles += (T + TC_NF - 1) / TC_NF;
  }
  P.meta[3 * B] = rows; P.meta[4 * B + 1] = tiles;
  P.rows = rows; P.tiles = tiles; P.span = cur;
  std::vector<int32_t> tc(tiles + 1, 0);
  for (int i = 0; i < B; ++i)
    for (int64_t t = P.meta[3 * B + 1 + i]; t < P.meta[3 * B + 2 + i]; ++t) tc[t] = i;
  P.meta.resize(4 * B + 2 + (tiles + 2) / 2);
  memcpy(&P.meta[4 * B + 2], tc.data(), tiles * 4);
  P.x.resize(cur + 8);
  srand(seed);
  for (auto &v : P.x) {  // ~N(0, 0.1^2) + a tone + DC so that the spectrum has dynamic range
    double u = 0;
    for (int j = 0; j < 12; ++j) u += (double)rand() / RAND_MAX;
    v = (float)(0.1 * (u - 6.0));
  }
  for (size_t i = 0; i < P.x.size(); ++i) P.x[i] += (float)(0.3 * sin(2 * M_PI * 440.0 * i / 16000.0) + 0.05);
}

// float64 evaluation of one frame: fills Y (16 x 17 complex), X (257 complex), mel energies
static void ref_frame(const Problem &P, int cut, int64_t t, std::vector<double> &v, std::vector<double> &Xr, std::vector<double> &Xi,
                      std::vector<double> &E) {
  const DevPlan &p = P.p;
  const int B = (int)P.ns.size();
  const int64_t n = P.ns[cut], off = P.meta[cut];
  std::vector<double> f(p.L);
  double mu = 0;
  for (int j = 0; j < p.L; ++j) {
    int64_t i = t * p.S - p.pad_left + j;
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
    f[j] = P.x[off + i];
    mu += f[j];
  }
  (void)B;
  mu /= p.L;
  v.assign(512, 0.0);
  for (int j = 0; j < p.L; ++j) {
    const double d = f[j] - mu, dp = f[j > 0 ? j - 1 : 0] - mu;
    v[j] = (d - (double)p.preemph * dp) * (double)P.window[j];
  }
  Xr.assign(257, 0.0); Xi.assign(257, 0.0);
  for (int k = 0; k <= 256; ++k) {
    double a = 0, b = 0;
    for (int j = 0; j < p.L; ++j) {
      const double ang = -2.0 * M_PI * (double)((int64_t)j * k % 512) / 512.0;
      a += v[j] * cos(ang); b += v[j] * sin(ang);
    }
    Xr[k] = a; Xi[k] = b;
  }
  E.assign(p.M, 0.0);
  for (int m = 0; m < p.M; ++m) {
    double e = 0;
    for (int k = 0; k < 257; ++k) e += (double)P.bank[(size_t)k * p.M + m] * (Xr[k] * Xr[k] + Xi[k] * Xi[k]);
    E[m] = e;
  }
}

struct DevProblem {
  float *x = nullptr, *out = nullptr, *dbg = nullptr;
  int64_t *meta = nullptr;
  Tc512Host hst;
  std::vector<void *> allocs;
  DevBatch db;
};

static void upload_problem(Problem &P, DevProblem &D) {
  const int B = (int)P.ns.size();
  CK(cudaMalloc(&D.x, P.x.size() * 4)); CK(cudaMemcpy(D.x, P.x.data(), P.x.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&D.meta, P.meta.size() * 8)); CK(cudaMemcpy(D.meta, P.meta.data(), P.meta.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&D.out, (size_t)P.rows * P.p.F * 4)); CK(cudaMemset(D.out, 0xff, (size_t)P.rows * P.p.F * 4));
  CK(cudaMalloc(&D.dbg, (2 * 128 * 32 + TC_NF * TC_PP) * 4)); CK(cudaMemset(D.dbg, 0, (2 * 128 * 32 + TC_NF * TC_PP) * 4));
  int fpt = 0;
  int rc = tc512_prepare(P.p, P.bank, D.allocs, &fpt, P.window, &D.hst);
  if (rc) { printf("tc512_prepare failed: %d\n", rc); exit(2); }
  DevBatch &db = D.db;
  memset(&db, 0, sizeof(db));
  db.samples = D.x; db.samp_off = D.meta; db.nsamp = D.meta + B; db.row_off = D.meta + 2 * B; db.tile_off = D.meta + 3 * B + 1;
  db.tile_cut = reinterpret_cast<const int32_t *>(D.meta + 4 * B + 2);
  db.out = D.out; db.tile_base = 0; db.num_tiles = P.tiles; db.max_frames = 0; db.batch_first = 0; db.B = B;
  db.out_mode = B200FEAT_OUT_PACKED; db.pad_value = 0.f;
}

static int tc_run() {
  Problem P;
  make_problem(P, {16000, 4000, 1599, 160000, 159, 2720}, 80, 1);
  DevProblem D;
  upload_problem(P, D);
  printf("tc: %lld rows, %lld tiles, smem %zu B, blob %d B\n", (long long)P.rows, (long long)P.tiles, D.hst.smem, D.hst.t.cblob_bytes);
  CK(cudaFuncSetAttribute(b200feat_tc512_kernel<B200FEAT_F32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.hst.smem));
  {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, b200feat_tc512_kernel<B200FEAT_F32, 0>));
    printf("tc: regs %d, local %zu B, smem %zu B\n", fa.numRegs, fa.localSizeBytes, D.hst.smem);
  }
  b200feat_tc512_kernel<B200FEAT_F32, 1><<<4, TC_THREADS, D.hst.smem>>>(P.p, D.hst.t, D.db, D.dbg);
  cudaError_t e = cudaDeviceSynchronize();
  printf("tc kernel status: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> dbg(2 * 128 * 32 + TC_NF * TC_PP), out((size_t)P.rows * 80);
  CK(cudaMemcpy(dbg.data(), D.dbg, dbg.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out.data(), D.out, out.size() * 4, cudaMemcpyDeviceToHost));
  // ---- stage checks on the first tile (cut 0, frames 0..7)
  std::vector<double> v, Xr, Xi, E;
  double eY = 0, mY = 0, eX = 0, mX = 0, eP = 0, mP = 0;
  for (int f = 0; f < TC_NF; ++f) {
    ref_frame(P, 0, f, v, Xr, Xi, E);
    for (int n2 = 0; n2 < 16; ++n2)
      for (int k1 = 0; k1 <= 16; ++k1) {
        double yr = 0, yi = 0;
        for (int n1 = 0; n1 < 32; ++n1) {
          const double a = -2.0 * M_PI * (double)((n1 * k1) % 32) / 32.0;
          yr += v[16 * n1 + n2] * cos(a); yi += v[16 * n1 + n2] * sin(a);
        }
        const float *d = &dbg[(f * 16 + n2) * 32];
        const double gr = k1 == 0 ? d[0] : k1 == 16 ? d[1] : d[2 * k1], gi = (k1 == 0 || k1 == 16) ? 0.0 : d[2 * k1 + 1];
        eY = fmax(eY, fmax(fabs(gr - yr), fabs(gi - yi))); mY = fmax(mY, fmax(fabs(yr), fabs(yi)));
      }
    for (int k1 = 0; k1 < 16; ++k1)
      for (int k2 = 0; k2 < 16; ++k2) {
        const int k = k1 + 32 * k2;
        if (k1 == 0 && k2 > 8) continue;
        const double xr = k <= 256 ? Xr[k] : Xr[512 - k], xi = k <= 256 ? Xi[k] : -Xi[512 - k];
        const float *d = &dbg[128 * 32 + (f * 16 + k1) * 32];
        eX = fmax(eX, fmax(fabs(d[2 * k2] - xr), fabs(d[2 * k2 + 1] - xi))); mX = fmax(mX, fmax(fabs(xr), fabs(xi)));
      }
    for (int k = 0; k <= 256; ++k) {
      const double pr = Xr[k] * Xr[k] + Xi[k] * Xi[k];
      eP = fmax(eP, fabs(dbg[2 * 128 * 32 + f * TC_PP + k] - pr)); mP = fmax(mP, pr);
    }
  }
  printf("stage 1 (D1 = Y):  max err %.3e  (max |Y| %.3e, rel %.2e)\n", eY, mY, eY / mY);
  printf("stage 2 (D2 = X):  max err %.3e  (max |X| %.3e, rel %.2e)\n", eX, mX, eX / mX);
  printf("power   (P)     :  max err %.3e  (max P %.3e, rel %.2e)\n", eP, mP, eP / mP);
  // ---- every output row
  const int B = (int)P.ns.size();
  double eO = 0; int64_t bad = 0, nanc = 0;
  for (int c = 0; c < B; ++c) {
    const int64_t T = (P.ns[c] + 80) / 160, r0 = P.meta[2 * B + c];
    double ec = 0;
    for (int64_t t = 0; t < T; ++t) {
      ref_frame(P, c, t, v, Xr, Xi, E);
      for (int m = 0; m < 80; ++m) {
        const double ref = log(fmax(E[m], (double)P.p.mel_floor));
        const float g = out[(r0 + t) * 80 + m];
        if (g != g) { ++nanc; continue; }
        const double er = fabs(g - ref);
        ec = fmax(ec, er);
        if (er > 2e-4 + 1e-4 * fabs(ref)) ++bad;
      }
    }
    printf("cut %d (n = %lld, T = %lld): max |log-mel err| %.3e\n", c, (long long)P.ns[c], (long long)T, ec);
    eO = fmax(eO, ec);
  }
  printf("output: max err %.3e, %lld values over the 2e-4 + 1e-4|ref| gate, %lld NaN\n", eO, (long long)bad, (long long)nanc);
  printf("%s\n", (bad == 0 && nanc == 0) ? "TC PROBE OK" : "TC PROBE MISMATCH");
  return (bad == 0 && nanc == 0) ? 0 : 1;
}

static int roles_main(int ncuts) {
  Problem P;
  make_problem(P, std::vector<int64_t>(ncuts, 160000), 80, 2);
  DevProblem D;
  upload_problem(P, D);
  CK(cudaFuncSetAttribute(b200feat_tc512_kernel<B200FEAT_F32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.hst.smem));
  CK(cudaMemset(D.dbg, 0, 128));
  b200feat_tc512_kernel<B200FEAT_F32, 2><<<148, TC_THREADS, D.hst.smem>>>(P.p, D.hst.t, D.db, D.dbg);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("roles: %s\n", cudaGetErrorString(e)); return 1; }
  unsigned long long h[16];
  CK(cudaMemcpy(h, D.dbg, 128, cudaMemcpyDeviceToHost));
  const char *nm[5] = {"INTER", "POWER", "PRE", "MEL", "ISSUE"};
  printf("roles (%llu tiles of 8 frames): cycles per tile", h[15]);
  for (int i = 0; i < 5; ++i) printf("  %s loop %.0f (waiting %.0f, working %.0f)", nm[i], (double)h[2 * i + 1] / h[15], (double)h[2 * i] / h[15], (double)(h[2 * i + 1] - h[2 * i]) / h[15]);
  printf("\n");
  return 0;
}

static int bench_main(int ncuts) {
  Problem P;
  make_problem(P, std::vector<int64_t>(ncuts, 160000), 80, 2);
  DevProblem D;
  upload_problem(P, D);
  int dev_sms = 148;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) tc512_launch(P.p, D.hst, D.db, B200FEAT_F32, dev_sms, 0);
  { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("bench: %s\n", cudaGetErrorString(e)); return 1; } }
  CK(cudaDeviceSynchronize());
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) tc512_launch(P.p, D.hst, D.db, B200FEAT_F32, dev_sms, 0);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  const double hours = (double)ncuts * 10.0 / 3600.0;
  printf("bench tc512: %d cuts x 10 s: %.3f ms per launch -> %.0f h audio/s, %.1f GB/s algorithmic (960 B/frame)\n", ncuts, ms, hours / (ms * 1e-3),
         (double)P.rows * 960.0 / (ms * 1e-3) / 1e9);
  return 0;
}

int main(int argc, char **argv) {
  const char *mode = argc > 1 ? argv[1] : "tc";
  if (!strcmp(mode, "disc")) return disc_main();
  if (!strcmp(mode, "rate")) return rate_main(argc > 2 ? atoi(argv[2]) : 0);
  if (!strcmp(mode, "tc")) return tc_run();
  if (!strcmp(mode, "bench")) return bench_main(argc > 2 ? atoi(argv[2]) : 2048);
  if (!strcmp(mode, "roles")) return roles_main(argc > 2 ? atoi(argv[2]) : 512);
  printf("unknown mode %s\n", mode);
  return 2;
}
