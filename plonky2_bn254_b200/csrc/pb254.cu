// Unity translation unit of libpb254.so: kernels + host orchestration + C ABI (include/pb254.h).
// Built by plonky2_bn254_b200/build.py with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// (or, test-only, g++ -x c++ -DPB254_HOSTSIM for the host-simulation library under tests/hostsim/).
#include "../../include/pb254.h"
#include "compat.cuh"
std::atomic<unsigned long long> g_pb_launches{0};
#include "context.cuh"
#include "ntt.cuh"
#include <thread>
#include "merkle.cuh"
#include "tracegen.h"
#include "prover.cuh"
#include "verify.cuh"

struct pb254_proof {
  prover::ProofData data;
  std::vector<u64> results;  // n_inputs x L limbs (pb254_prove / pb254_prove_dev only)
};

namespace {
thread_local std::string g_last_error;

template <class F>
int guarded(F&& f) {
  try {
    f();
    return PB254_OK;
  } catch (const Pb254Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return PB254_E_BAD_ARG;
  }
}

int ilog2_strict(size_t n) {
  int k = 0;
  while (((size_t)1 << k) < n) k++;
  if (((size_t)1 << k) != n) throw Pb254Error(PB254_E_BAD_ARG, "size is not a power of two");
  return k;
}

// One argument check for every extern "C" entry: never dereference before these pass (PB254_E_BAD_ARG).
void need(bool ok, const char* what) {
  if (!ok) throw Pb254Error(PB254_E_BAD_ARG, what);
}
void need_ctx(const pb254_ctx* c) { need(c != nullptr, "null context"); }
void need_kind(int kind) { need(kind >= 0 && kind <= 2, "unknown STARK kind (0 = G1, 1 = G2, 2 = Fq)"); }
// trace heights: a power of two, at least 2^16 (the range-check table needs 65536 rows) and at most 2^26
int need_trace_rows(size_t n_rows) {
  int k = 0;
  while (k < 27 && ((size_t)1 << k) < n_rows) k++;
  need(((size_t)1 << k) == n_rows && k >= 16 && k <= 26, "trace height must be a power of two in 2^16 .. 2^26");
  return k;
}
pb254_config config_or_default(const pb254_config* cfg_in, int log_rows) {
  pb254_config cfg;
  if (cfg_in)
    cfg = *cfg_in;
  else
    pb254_config_standard_fast(&cfg);
  prover::validate_config(cfg);
  if (log_rows >= 0) need(cfg.cap_height <= (unsigned)log_rows + cfg.rate_bits, "cap_height larger than log2 of the LDE size");
  return cfg;
}

struct Shape {
  int L, width, aux_len, in_words;
};
Shape shape_for(int kind) {
  switch (kind) {
    case PB254_KIND_G1: return {32, 781, 354, 20};
    case PB254_KIND_G2: return {64, 1295, 708, 36};
    case PB254_KIND_FQ: return {16, 427, 80, 8};
  }
  throw Pb254Error(PB254_E_BAD_ARG, "unknown STARK kind");
}

void throw_trace_error(int herr) {
  if (herr == tg::ERR_NOT_ON_CURVE)
    throw Pb254Error(PB254_E_NOT_ON_CURVE, "an input point is not on the curve (G1: y^2 = x^3 + 3, G2: y^2 = x^3 + 3/(9+u))");
  if (herr == tg::ERR_NOT_CANONICAL) throw Pb254Error(PB254_E_NOT_CANONICAL, "input coordinate >= p");
  if (herr == tg::ERR_INFINITY)
    throw Pb254Error(PB254_E_INFINITY, "an intermediate sum is the point at infinity (a = -b), unsupported by design");
  if (herr) throw Pb254Error(PB254_E_BAD_ARG, "trace generation: witness consistency check failed");
}

// results[k][i] = limb i of the `sum` / `product` register at the last row of instance k: s * x + offset (x^s)
struct GatherResultsK {
  const u64* trace;
  u64* out;
  size_t n_rows;
  int reg1, L;
  PB_HD void operator()(size_t gid) const {
    size_t k = gid / L;
    int i = (int)(gid % L);
    out[gid] = trace[(size_t)(reg1 + i) * n_rows + k * tg::PERIOD + (tg::PERIOD - 1)];
  }
};

struct PermuteK {
  const u64* in;
  u64* out;
  PB_HD void operator()(size_t i) const {
    u64 s[12];
    for (int k = 0; k < 12; k++) s[k] = in[i * 12 + k];
    poseidon::permute(s);
    for (int k = 0; k < 12; k++) out[i * 12 + k] = s[k];
  }
};
}  // namespace

extern "C" {

void pb254_config_standard_fast(pb254_config* c) {
  c->rate_bits = 1;
  c->cap_height = 4;
  c->num_challenges = 2;
  c->num_query_rounds = 84;
  c->pow_bits = 16;
  c->arity_bits = 4;
  c->final_poly_bits = 5;
}

const char* pb254_last_error(void) { return g_last_error.c_str(); }
uint64_t pb254_launch_count(void) { return g_pb_launches.load(); }

int pb254_ctx_create(int device, void* stream, pb254_ctx** out) {
  return guarded([&] {
    if (!out) throw Pb254Error(PB254_E_BAD_ARG, "null out pointer");
    PbDeviceGuard device_guard(device);
    pb254_ctx* c = new pb254_ctx();
    c->device = device;
#if PB_HOSTSIM
    c->stream = stream;
#else
    if (stream) {
      c->stream = (cudaStream_t)stream;
    } else {
      PB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      c->own_stream = true;
    }
#endif
    c->tables.init(c->stream);
    *out = c;
  });
}

void pb254_ctx_destroy(pb254_ctx* c) {
  if (!c) return;
  c->times.clear();
  c->tables.destroy();
  c->arena.destroy();
#if !PB_HOSTSIM
  if (c->own_stream) cudaStreamDestroy(c->stream);
#endif
  delete c;
}

/* stage timings of the last call on this context (device time, ms) */
int pb254_timing_count(pb254_ctx* c) { return c ? (int)c->times.recs.size() : 0; }
const char* pb254_timing_name(pb254_ctx* c, int i) {
  return c && i >= 0 && i < (int)c->times.recs.size() ? c->times.recs[i].name.c_str() : "";
}
double pb254_timing_ms(pb254_ctx* c, int i) {
  return c && i >= 0 && i < (int)c->times.recs.size() ? c->times.recs[i].ms : 0.0;
}

int pb254_trace_width(int kind) { return kind >= 0 && kind <= 2 ? shape_for(kind).width : -1; }
int pb254_input_words(int kind) { return kind >= 0 && kind <= 2 ? shape_for(kind).in_words : -1; }
int pb254_num_aux(int kind, uint32_t nch) {
  if (kind < 0 || kind > 2) return -1;
  Shape s = shape_for(kind);
  int ncols = 3 * s.L + s.aux_len;
  return ((ncols + 1) / 2 + 1 + 2) * (int)nch;
}
size_t pb254_trace_rows(size_t n_inputs, size_t min_rows) {
  size_t n = n_inputs * 512 > min_rows ? n_inputs * 512 : min_rows, r = 1;
  while (r < n) r <<= 1;
  return r;
}

int pb254_poseidon_permute(pb254_ctx* c, const uint64_t* in, size_t n, uint64_t* out) {
  return guarded([&] {
    need_ctx(c);
    PbDeviceGuard device_guard(c->device);
    c->arena.reserve(2 * n * 96 + 1024);
    c->arena.reset();
    u64* din = c->arena.alloc_n<u64>(n * 12);
    u64* dout = c->arena.alloc_n<u64>(n * 12);
    pb_h2d(din, in, n * 96, c->stream);
    pb_launch("poseidon permute", PermuteK{din, dout}, n, c->stream, 128);
    pb_d2h(out, dout, n * 96, c->stream);
    pb_sync(c->stream);
  });
}

int pb254_lde_batch(pb254_ctx* c, const uint64_t* values, size_t cols, size_t n, uint32_t rate_bits, int from_coeffs,
                    uint64_t* lde_out) {
  return guarded([&] {
    need_ctx(c);
    PbDeviceGuard device_guard(c->device);
    int L = ilog2_strict(n);
    size_t N = n << rate_bits;
    c->arena.reserve((2 * cols * n + cols * N) * 8 + 4096);
    c->arena.reset();
    u64* dv = c->arena.alloc_n<u64>(cols * n);
    u64* scratch = c->arena.alloc_n<u64>(cols * n);
    u64* lde = c->arena.alloc_n<u64>(cols * N);
    pb_h2d(dv, values, cols * n * 8, c->stream);
    ntt::lde_columns(c->tables, dv, n, lde, N, scratch, (int)cols, L, (int)rate_bits,
                     from_coeffs ? ntt::FROM_COEFFS_LDE : ntt::FROM_VALUES_LDE, c->stream);
    pb_d2h(lde_out, lde, cols * N * 8, c->stream);
    pb_sync(c->stream);
  });
}

int pb254_commit(pb254_ctx* c, const uint64_t* values, size_t cols, size_t n, uint32_t rate_bits, uint32_t cap_height,
                 int from_coeffs, uint64_t* cap_out, uint64_t* digests_out) {
  return guarded([&] {
    need_ctx(c);
    PbDeviceGuard device_guard(c->device);
    int L = ilog2_strict(n);
    size_t N = n << rate_bits;
    int log_N = L + (int)rate_bits;
    size_t nd = merkle::tree_digests(log_N, (int)cap_height);
    c->arena.reserve((2 * cols * n + cols * N) * 8 + nd * 32 + 4096);
    c->arena.reset();
    u64* dv = c->arena.alloc_n<u64>(cols * n);
    u64* scratch = c->arena.alloc_n<u64>(cols * n);
    u64* lde = c->arena.alloc_n<u64>(cols * N);
    merkle::Digest* dig = c->arena.alloc_n<merkle::Digest>(nd);
    pb_h2d(dv, values, cols * n * 8, c->stream);
    c->times.clear();
    int t0 = c->times.begin("lde", c->stream);
    ntt::lde_columns(c->tables, dv, n, lde, N, scratch, (int)cols, L, (int)rate_bits,
                     from_coeffs ? ntt::FROM_COEFFS_LDE : ntt::FROM_VALUES_LDE, c->stream);
    c->times.end(t0, c->stream);
    int t1 = c->times.begin("merkle", c->stream);
    merkle::build_from_lde(lde, N, (int)cols, log_N, (int)cap_height, dig, c->stream);
    c->times.end(t1, c->stream);
    size_t ncap = (size_t)1 << cap_height;
    pb_d2h(cap_out, dig + (nd - ncap), ncap * 32, c->stream);
    if (digests_out) pb_d2h(digests_out, dig, nd * 32, c->stream);
    pb_sync(c->stream);
    c->times.resolve();
  });
}

int pb254_generate_trace(pb254_ctx* c, int kind, const uint64_t* inputs, const uint64_t* timestamps, size_t n_inputs,
                         size_t min_rows, uint64_t* cols_out) {
  return guarded([&] {
    need_ctx(c);
    need_kind(kind);
    need(inputs && timestamps && cols_out && n_inputs > 0, "null or empty argument");
    PbDeviceGuard device_guard(c->device);
    tg::Layout l = tg::layout_for(kind);
    size_t n_rows = pb254_trace_rows(n_inputs, min_rows);
    need_trace_rows(n_rows);
    size_t tbytes = (size_t)l.width * n_rows * 8;
    c->arena.reserve(tbytes + tg::scratch_bytes(kind, n_inputs) + n_inputs * (l.in_words + 1) * 8 + 65536);
    c->arena.reset();
    u64* d_trace = c->arena.alloc_n<u64>((size_t)l.width * n_rows);
    u64* d_in = c->arena.alloc_n<u64>(n_inputs * l.in_words + 1);
    u64* d_ts = c->arena.alloc_n<u64>(n_inputs + 1);
    int* d_err = c->arena.alloc_n<int>(1);
    pb_h2d(d_in, inputs, n_inputs * l.in_words * 8, c->stream);
    pb_h2d(d_ts, timestamps, n_inputs * 8, c->stream);
    pb_memset(d_err, 0, sizeof(int), c->stream);
    c->times.clear();
    int t0 = c->times.begin("tracegen", c->stream);
    tg::generate(c->arena, kind, d_in, d_ts, n_inputs, n_rows, d_trace, d_err, c->stream);
    c->times.end(t0, c->stream);
    int herr = 0;
    pb_d2h(&herr, d_err, sizeof(int), c->stream);
    pb_d2h(cols_out, d_trace, tbytes, c->stream);
    pb_sync(c->stream);
    c->times.resolve();
    throw_trace_error(herr);
  });
}

// generate_trace + prove in one call: what run_once does between
// src/generators/g1/stark_proof.rs:154 and :163. The trace never leaves the device.
static int prove_inputs_impl(pb254_ctx* c, int kind, const uint64_t* inputs, const uint64_t* timestamps, size_t n_inputs,
                             size_t min_rows, const pb254_config* cfg_in, int keep_debug, pb254_proof** out,
                             bool inputs_on_device, const pb254_comm* comm = nullptr) {
  return guarded([&] {
    need(out != nullptr, "null out pointer");
    need_ctx(c);
    need_kind(kind);
    need(inputs && timestamps && n_inputs > 0, "null or empty input");
    PbDeviceGuard device_guard(c->device);
    tg::Layout l = tg::layout_for(kind);
    size_t n_rows = pb254_trace_rows(n_inputs, min_rows);
    const pb254_config cfg = config_or_default(cfg_in, need_trace_rows(n_rows));
    size_t twords = (size_t)l.width * n_rows;
    size_t tg_bytes = tg::scratch_bytes(kind, n_inputs) + n_inputs * (l.in_words + 1) * 8 + 65536;
    need(!comm || (comm->world >= 1 && (comm->world & (comm->world - 1)) == 0 && comm->rank < comm->world),
         "comm: world must be a power of two and rank < world");
    size_t pv_bytes = comm && comm->world > 1 ? prover::workspace_bytes_sharded(kind, n_rows, cfg, comm->world)
                                              : prover::workspace_bytes(kind, n_rows, cfg);
    c->arena.reserve(twords * 8 + (tg_bytes > pv_bytes ? tg_bytes : pv_bytes) + 65536);
    c->arena.reset();
    c->times.clear();
    u64* d_trace = c->arena.alloc_n<u64>(twords);
    size_t mark = c->arena.off;
    std::vector<u64> results;
    {
      prover::Stage st(c, "tracegen");
      const u64 *d_in = inputs, *d_ts = timestamps;
      if (!inputs_on_device) {
        u64* hin = c->arena.alloc_n<u64>(n_inputs * l.in_words + 1);
        u64* hts = c->arena.alloc_n<u64>(n_inputs + 1);
        pb_h2d(hin, inputs, n_inputs * l.in_words * 8, c->stream);
        pb_h2d(hts, timestamps, n_inputs * 8, c->stream);
        d_in = hin;
        d_ts = hts;
      }
      int* d_err = c->arena.alloc_n<int>(1);
      pb_memset(d_err, 0, sizeof(int), c->stream);
      tg::generate(c->arena, kind, d_in, d_ts, n_inputs, n_rows, d_trace, d_err, c->stream);
      int herr = 0;
      pb_d2h(&herr, d_err, sizeof(int), c->stream);
      u64* d_res = c->arena.alloc_n<u64>(n_inputs * l.L);
      pb_launch("gather results", GatherResultsK{d_trace, d_res, n_rows, l.reg1, l.L}, n_inputs * l.L, c->stream, 128);
      results.resize(n_inputs * l.L);
      pb_d2h(results.data(), d_res, results.size() * 8, c->stream);
      pb_sync(c->stream);
      throw_trace_error(herr);
    }
    c->arena.off = mark;
    pb254_proof* pf = new pb254_proof();
    pf->results.swap(results);
    try {
      prover::prove_device(c, kind, d_trace, n_rows, cfg, pf->data, keep_debug != 0, comm);
    } catch (...) {
      delete pf;
      throw;
    }
    pb_sync(c->stream);
    c->times.resolve();
    *out = pf;
  });
}

// One proof across the ranks of `comm` (include/pb254.h). When the trace splits into per-rank row blocks of whole
// instances that each hold the 2^16-row range-check table or none of it (n / world a multiple of 512 and >= 65536 -
// every BASELINE size does), every rank generates only ITS instances: the row block [rank n / P, (rank + 1) n / P) of
// the trace; the range-check histogram is added up across the ranks and lands in rank 0's frequency column. Otherwise
// (small traces) every rank generates the whole trace. world == 1 (or comm == NULL) is pb254_prove.
int pb254_prove_sharded(pb254_ctx* c, int kind, const uint64_t* inputs, const uint64_t* timestamps, size_t n_inputs,
                        size_t min_rows, const pb254_config* cfg_in, const pb254_comm* comm, pb254_proof** out) {
  if (!comm || comm->world <= 1)
    return prove_inputs_impl(c, kind, inputs, timestamps, n_inputs, min_rows, cfg_in, 0, out, false, comm);
  {
    const size_t n_rows = pb254_trace_rows(n_inputs, min_rows), P = comm->world;
    const bool pow2 = P && (P & (P - 1)) == 0;
    if (!pow2 || n_rows % P || (n_rows / P) % tg::PERIOD || n_rows / P < 65536)
      return prove_inputs_impl(c, kind, inputs, timestamps, n_inputs, min_rows, cfg_in, 0, out, false, comm);
  }
  return guarded([&] {
    need(out != nullptr, "null out pointer");
    need_ctx(c);
    need_kind(kind);
    need(inputs && timestamps && n_inputs > 0, "null or empty input");
    need(comm->rank < comm->world && comm->all_to_all && comm->all_gather, "comm: rank < world, callbacks non-null");
    PbDeviceGuard device_guard(c->device);
    const tg::Layout l = tg::layout_for(kind);
    const size_t n_rows = pb254_trace_rows(n_inputs, min_rows), P = comm->world, rk = comm->rank, nloc = n_rows / P;
    const pb254_config cfg = config_or_default(cfg_in, need_trace_rows(n_rows));
    const size_t per = nloc / tg::PERIOD;  // instances per rank
    const size_t i0 = std::min(n_inputs, rk * per), i1 = std::min(n_inputs, (rk + 1) * per), kloc = i1 - i0;
    const size_t Wpad = ((size_t)l.width + P - 1) / P * P;
    const size_t twords = Wpad * nloc;
    size_t tg_bytes = tg::scratch_bytes(kind, kloc ? kloc : 1) + (kloc + 1) * (l.in_words + 1) * 8 + (P + 1) * 65536 * 8 + 65536;
    size_t pv_bytes = prover::workspace_bytes_sharded(kind, n_rows, cfg, P);
    size_t want = twords * 8 + (tg_bytes > pv_bytes ? tg_bytes : pv_bytes) + 65536;
#if !PB_HOSTSIM
    {  // the estimate is an upper bound: never ask for more than the device has left (the collectives of the caller
       // need room too); a proof that really does not fit fails in Arena::alloc with PB254_E_OOM
      size_t free_b = 0, total_b = 0;
      PB_CUDA(cudaMemGetInfo(&free_b, &total_b));
      const size_t avail = free_b + c->arena.cap, margin = (size_t)4 << 30;
      if (avail > margin && want > avail - margin) want = avail - margin;
    }
#endif
    c->arena.reserve(want);
    c->arena.reset();
    c->times.clear();
    u64* d_rows = c->arena.alloc_n<u64>(twords);
    size_t mark = c->arena.off;
    {
      prover::Stage st(c, "tracegen");
      u64* d_in = c->arena.alloc_n<u64>(kloc * l.in_words + 1);
      u64* d_ts = c->arena.alloc_n<u64>(kloc + 1);
      if (kloc) {
        pb_h2d(d_in, inputs + i0 * l.in_words, kloc * l.in_words * 8, c->stream);
        pb_h2d(d_ts, timestamps + i0, kloc * 8, c->stream);
      }
      int* d_err = c->arena.alloc_n<int>(1);
      pb_memset(d_err, 0, sizeof(int), c->stream);
      tg::generate(c->arena, kind, d_in, d_ts, kloc, nloc, d_rows, d_err, c->stream, rk * nloc);
      // range-check frequencies: every block counted its own cells into the first 2^16 rows of its frequency column;
      // the table rows are rows 0 .. 65535 of the whole trace, i.e. of rank 0's block
      u64* freq = d_rows + (size_t)l.freq * nloc;
      u64* all = c->arena.alloc_n<u64>(P * 65536);
      if (comm->all_gather(comm->user, freq, all, 65536 * 8)) throw Pb254Error(PB254_E_CUDA, "collective failed: all_gather");
      if (rk == 0)
        pb_launch("sum histograms", prover::SumRecordsK{all, freq, 65536, (int)P}, 65536, c->stream, 128);
      else
        pb_memset(freq, 0, 65536 * 8, c->stream);
      // an input error on one rank must fail the call on every rank (the others would wait in the next collective)
      u64* e_mine = c->arena.alloc_n<u64>(1);
      u64* e_all = c->arena.alloc_n<u64>(P);
      pb_memset(e_mine, 0, 8, c->stream);
      pb_d2d(e_mine, d_err, sizeof(int), c->stream);
      if (comm->all_gather(comm->user, e_mine, e_all, 8)) throw Pb254Error(PB254_E_CUDA, "collective failed: all_gather");
      std::vector<u64> herr(P);
      pb_d2h(herr.data(), e_all, P * 8, c->stream);
      pb_sync(c->stream);
      for (size_t p = 0; p < P; p++) throw_trace_error((int)(herr[p] & 0xffffffffu));
    }
    c->arena.off = mark;
    pb254_proof* pf = new pb254_proof();  // no native results in this mode: they are spread over the ranks
    try {
      prover::prove_device(c, kind, d_rows, n_rows, cfg, pf->data, false, comm, true);
    } catch (...) {
      delete pf;
      throw;
    }
    pb_sync(c->stream);
    c->times.resolve();
    *out = pf;
  });
}

int pb254_prove(pb254_ctx* c, int kind, const uint64_t* inputs, const uint64_t* timestamps, size_t n_inputs,
                size_t min_rows, const pb254_config* cfg, int keep_debug, pb254_proof** out) {
  return prove_inputs_impl(c, kind, inputs, timestamps, n_inputs, min_rows, cfg, keep_debug, out, false);
}

int pb254_prove_many(pb254_ctx* const* ctxs, size_t n_ctx, int kind, const uint64_t* inputs, const uint64_t* timestamps,
                     size_t n_inputs, size_t n_batches, size_t min_rows, const pb254_config* cfg, pb254_proof** proofs_out) {
  int rc = guarded([&] {
    need(ctxs && n_ctx > 0 && n_ctx <= 16 && proofs_out, "null contexts / outputs or more than 16 contexts");
    need_kind(kind);
    need(inputs && timestamps && n_inputs > 0, "null or empty input");
    for (size_t t = 0; t < n_ctx; t++) {
      need_ctx(ctxs[t]);
      for (size_t u = 0; u < t; u++) need(ctxs[u] != ctxs[t], "the same context twice");
    }
  });
  if (rc) return rc;
  const size_t in_words = (size_t)shape_for(kind).in_words;
  for (size_t b = 0; b < n_batches; b++) proofs_out[b] = nullptr;
  std::vector<int> codes(n_ctx, 0);
  std::vector<std::string> messages(n_ctx);
  std::vector<std::thread> workers;
  for (size_t t = 0; t < n_ctx; t++)
    workers.emplace_back([&, t] {
      for (size_t b = t; b < n_batches && !codes[t]; b += n_ctx) {
        codes[t] = pb254_prove(ctxs[t], kind, inputs + b * n_inputs * in_words, timestamps + b * n_inputs, n_inputs,
                               min_rows, cfg, 0, &proofs_out[b]);
        if (codes[t]) messages[t] = pb254_last_error();  // thread-local in this worker
      }
    });
  for (auto& w : workers) w.join();
  for (size_t t = 0; t < n_ctx; t++)
    if (codes[t]) {
      for (size_t b = 0; b < n_batches; b++) {
        if (proofs_out[b]) pb254_proof_free(proofs_out[b]);
        proofs_out[b] = nullptr;
      }
      g_last_error = messages[t];
      return codes[t];
    }
  return PB254_OK;
}

// Same with `inputs` / `timestamps` already resident in device memory of the context's GPU.
int pb254_prove_dev(pb254_ctx* c, int kind, const uint64_t* d_inputs, const uint64_t* d_timestamps, size_t n_inputs,
                    size_t min_rows, const pb254_config* cfg, int keep_debug, pb254_proof** out) {
  return prove_inputs_impl(c, kind, d_inputs, d_timestamps, n_inputs, min_rows, cfg, keep_debug, out, true);
}

// prove(stark, config, trace, ...) on a host trace (column-major width x n_rows), the literal
// signature of src/starks/common/prover.rs:18-30.
int pb254_prove_trace(pb254_ctx* c, int kind, const uint64_t* trace_cols, size_t n_rows, const pb254_config* cfg_in,
                      int keep_debug, pb254_proof** out) {
  return guarded([&] {
    need(out != nullptr, "null out pointer");
    need_ctx(c);
    need_kind(kind);
    need(trace_cols != nullptr, "null trace");
    PbDeviceGuard device_guard(c->device);
    const pb254_config cfg = config_or_default(cfg_in, need_trace_rows(n_rows));
    tg::Layout l = tg::layout_for(kind);
    size_t twords = (size_t)l.width * n_rows;
    c->arena.reserve(twords * 8 + prover::workspace_bytes(kind, n_rows, cfg) + 65536);
    c->arena.reset();
    c->times.clear();
    u64* d_trace = c->arena.alloc_n<u64>(twords);
    pb_h2d(d_trace, trace_cols, twords * 8, c->stream);
    {
      // A host-supplied trace has not been through generate_range_checks: every looked-up cell and the table column
      // must be < 2^16 (the reference asserts this while filling the frequency column, g1/scalar_mul_stark.rs:71-87).
      size_t mark = c->arena.off;
      int* d_err = c->arena.alloc_n<int>(1);
      pb_memset(d_err, 0, sizeof(int), c->stream);
      pb_launch("range pre-check", aux::RangePrecheckK{d_trace, n_rows, l.rc_lo, l.rc_hi, l.range_counter, d_err},
                n_rows, c->stream, 128);
      int herr = 0;
      pb_d2h(&herr, d_err, sizeof(int), c->stream);
      pb_sync(c->stream);
      c->arena.off = mark;
      if (herr) throw Pb254Error(PB254_E_BAD_ARG, "trace: a range-checked cell or the range counter is >= 2^16");
    }
    pb254_proof* pf = new pb254_proof();
    try {
      prover::prove_device(c, kind, d_trace, n_rows, cfg, pf->data, keep_debug != 0);
    } catch (...) {
      delete pf;
      throw;
    }
    pb_sync(c->stream);
    c->times.resolve();
    *out = pf;
  });
}

// ---- device-pointer building blocks of the oversized-trace mode (column-sharded LDE -> all-to-all ->
// row-sharded leaf hashing -> digest all-gather -> per-rank subtree), orchestrated by plonky2_bn254_b200/dist.py
int pb254_lde_dev(pb254_ctx* c, const uint64_t* d_values, size_t cols, size_t n, uint32_t rate_bits, int from_coeffs,
                  uint64_t* d_lde_out) {
  return guarded([&] {
    need_ctx(c);
    PbDeviceGuard device_guard(c->device);
    int L = ilog2_strict(n);
    c->arena.reserve(cols * n * 8 + 4096);
    c->arena.reset();
    u64* scratch = c->arena.alloc_n<u64>(cols * n);
    c->times.clear();
    int t0 = c->times.begin("lde", c->stream);
    ntt::lde_columns(c->tables, d_values, n, d_lde_out, n << rate_bits, scratch, (int)cols, L, (int)rate_bits,
                     from_coeffs ? ntt::FROM_COEFFS_LDE : ntt::FROM_VALUES_LDE, c->stream);
    c->times.end(t0, c->stream);
    pb_sync(c->stream);
    c->times.resolve();
  });
}

int pb254_leaf_hash_rows_dev(pb254_ctx* c, const uint64_t* d_matrix, size_t stride, size_t cols, size_t rows,
                             uint64_t* d_digests_out) {
  return guarded([&] {
    need_ctx(c);
    PbDeviceGuard device_guard(c->device);
    c->times.clear();
    int t0 = c->times.begin("leaf hash rows", c->stream);
    merkle::hash_rows_natural(d_matrix, stride, (int)cols, rows, (merkle::Digest*)d_digests_out, c->stream);
    c->times.end(t0, c->stream);
    pb_sync(c->stream);
    c->times.resolve();
  });
}

// Roots at height `log_sub - log_roots` of the subtree whose leaves are tree positions [first, first + 2^log_sub):
// leaf(q) = d_all_digests[bit_reverse(q, log_total)] (digests of all rows in natural row order).
int pb254_merkle_subtree_dev(pb254_ctx* c, const uint64_t* d_all_digests, uint32_t log_total, size_t first,
                             uint32_t log_sub, uint32_t log_roots, uint64_t* d_roots_out) {
  return guarded([&] {
    need_ctx(c);
    PbDeviceGuard device_guard(c->device);
    if (log_roots > log_sub || log_sub > log_total) throw Pb254Error(PB254_E_BAD_ARG, "subtree shape");
    size_t nd = merkle::tree_digests((int)log_sub, (int)log_roots);
    c->arena.reserve(nd * 32 + 4096);
    c->arena.reset();
    merkle::Digest* dig = c->arena.alloc_n<merkle::Digest>(nd);
    c->times.clear();
    int t0 = c->times.begin("subtree", c->stream);
    pb_launch("subtree leaves", merkle::SubtreeLeavesK{(const merkle::Digest*)d_all_digests, dig, first, (int)log_total},
              (size_t)1 << log_sub, c->stream, 128);
    merkle::build_levels(dig, (int)log_sub, (int)log_roots, c->stream);
    size_t nr = (size_t)1 << log_roots;
    pb_d2d(d_roots_out, dig + (nd - nr), nr * 32, c->stream);
    c->times.end(t0, c->stream);
    pb_sync(c->stream);
    c->times.resolve();
  });
}

// verify(stark, config, ctls, proof, [], extra_looking_values) of src/starks/common/verifier.rs:32-98, with the
// extra looking values recomputed natively from the batch as run_once does (g1/scalar_mul_ctl.rs:57-80).
int pb254_verify(int kind, const pb254_config* cfg_in, const uint64_t* proof_words, size_t n_words,
                 const uint64_t* inputs, const uint64_t* timestamps, size_t n_inputs) {
  return guarded([&] {
    need_kind(kind);
    need(proof_words && inputs && timestamps, "null argument");
    const pb254_config cfg = config_or_default(cfg_in, -1);
    verify::verify_proof(kind, cfg, proof_words, n_words, inputs, timestamps, n_inputs);
  });
}

void pb254_proof_free(pb254_proof* p) { delete p; }
size_t pb254_proof_results_words(const pb254_proof* p) { return p->results.size(); }
const uint64_t* pb254_proof_results_data(const pb254_proof* p) { return p->results.data(); }
int pb254_proof_parse(const uint64_t* proof_words, size_t n_words, pb254_proof_layout* out) {
  return guarded([&] {
    if (!proof_words || !out) throw Pb254Error(PB254_E_BAD_ARG, "null argument");
    proofview::parse(proof_words, n_words, *out);
  });
}

size_t pb254_proof_words(const pb254_proof* p) { return p->data.blob.size(); }
const uint64_t* pb254_proof_data(const pb254_proof* p) { return p->data.blob.data(); }
// debug artefacts kept when keep_debug != 0: 0 auxiliary values (A x n), 1 quotient chunk coefficients
// (2*num_challenges x n), 2 challenges [betas, gammas, alphas, zeta, fri_alpha, fri_betas], 3 query indices
size_t pb254_proof_debug_words(const pb254_proof* p, int which) {
  const std::vector<u64>* v = which == 0 ? &p->data.dbg_aux : which == 1 ? &p->data.dbg_chunks
                            : which == 2 ? &p->data.dbg_challenges : which == 3 ? &p->data.dbg_indices : which == 4 ? &p->data.dbg_qvals : which == 5 ? &p->data.dbg_groups : nullptr;
  return v ? v->size() : 0;
}
const uint64_t* pb254_proof_debug_data(const pb254_proof* p, int which) {
  const std::vector<u64>* v = which == 0 ? &p->data.dbg_aux : which == 1 ? &p->data.dbg_chunks
                            : which == 2 ? &p->data.dbg_challenges : which == 3 ? &p->data.dbg_indices : which == 4 ? &p->data.dbg_qvals : which == 5 ? &p->data.dbg_groups : nullptr;
  return v ? v->data() : nullptr;
}

}  // extern "C"
