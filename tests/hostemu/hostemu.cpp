/*
 * TEST-ONLY.  Single-lane host build of distributed_cluster_gpus_b200/csrc/dcsim_core.cuh
 * (DCSIM_LANES = 1; warp collectives are identities).  Lets the CPU test-suite run the *device source's*
 * handler logic, state-block layout, Philox window, FIFO rings and resume path against the oracle where no
 * GPU exists.  It is not part of, linked into, or reachable from the product library — the product fails
 * loudly without CUDA.  What it cannot show (lane mapping, __syncwarp placement, REDUX) is what the
 * `-m gpu` parity tests are for.
 */
#define DCSIM_HOST_EMU 1
#ifdef DCSIM_HOSTEMU_PERTURB
/* Conditioning probe (second build, libdcsim_hostemu_perturbed.so): every 5th pow() result is moved by ONE ulp — the
 * kind of difference two correct libm implementations (glibc here, CUDA's on the device) have.  How far a scenario's
 * results move under it is that scenario's sensitivity to last-bit differences: ~1e-15 for ordinary scenarios, orders
 * of magnitude more where the power-cap controller keeps re-timing back-to-back jobs (each re-timing multiplies a
 * start-time error by rate_old/rate_new).  tools/fuzz_gpu.py uses it to tell ill-conditioned scenarios from bugs. */
#include <math.h>
static inline double dcsim_hostemu_perturbed_pow(double a, double b) {
  static unsigned long calls = 0;
  const double r = pow(a, b);
  return (++calls % 5ul == 0ul) ? nextafter(r, INFINITY) : r;
}
#define pow dcsim_hostemu_perturbed_pow
#endif
#include "../../distributed_cluster_gpus_b200/csrc/dcsim_core.cuh"

#include <stdio.h>
#include <stdlib.h>

extern "C" {

size_t hostemu_sizeof_spec(void) { return sizeof(dcsim_spec_t); }
void hostemu_set_test_time_quantum(double q) { dcsim_test_time_quantum = q; } /* see dcsim_core.cuh dcsim_test_quantize */

/* Runs n replicas; each advance call processes `chunk_events` events per replica (0 = to the end) and the
 * state block round-trips through "HBM" between calls exactly as the kernel's stage-in/stage-out does.
 * trace/job/cluster recorders apply to `rec_replica` (-1 = none); counts[3] receives the row counts. */
long long hostemu_run_batch(const void* spec_blob, size_t spec_bytes, uint64_t n_replicas, uint64_t seed0,
                            uint64_t chunk_events, double* out_summaries, int64_t rec_replica,
                            dcsim_trace_rec_t* trace, uint32_t trace_cap, dcsim_job_rec_t* jobs, uint32_t jobs_cap,
                            dcsim_cluster_rec_t* cluster, uint32_t cluster_cap, uint32_t* counts, int32_t* layout_out,
                            uint32_t* lat_hist /* [n][2][DCSIM_LAT_BINS] or NULL */, int rng_kind /* 0 Philox, 1 MT19937 */) {
  if (!spec_blob || spec_bytes != sizeof(dcsim_spec_t)) return -1;
  dcsim_kparams_t* P = (dcsim_kparams_t*)calloc(1, sizeof(dcsim_kparams_t));
  memcpy(&P->spec, spec_blob, sizeof(dcsim_spec_t));
  if (P->spec.magic != DCSIM_SPEC_MAGIC) { free(P); return -1; }
  dcsim_make_layout(&P->spec, &P->L, /*job_log=*/jobs != NULL);
  P->cap_arr = (uint32_t)(P->spec.cap_arrivals > 0 ? P->spec.cap_arrivals : 16384);
  if (layout_out) { layout_out[0] = P->L.total_bytes; layout_out[1] = (P->L.xring_mask + 1) / 2; layout_out[2] = P->L.cap_run;
                    layout_out[3] = P->L.cap_q[0]; layout_out[4] = P->L.cap_q[1]; layout_out[5] = P->L.rec_off; }
  uint32_t local_counts[4] = {0, 0, 0, 0};
  P->rec.trace = trace; P->rec.jobs = jobs; P->rec.cluster = cluster; P->rec.counts = counts ? counts : local_counts;
  P->rec.trace_cap = trace_cap; P->rec.jobs_cap = jobs_cap; P->rec.cluster_cap = cluster_cap;
  P->rec.trace_replica = trace ? rec_replica : -1; P->rec.log_replica = (jobs || cluster) ? rec_replica : -1;
  if (counts) counts[0] = counts[1] = counts[2] = 0;
  P->n_replicas = n_replicas; P->seed0 = seed0; P->max_events = chunk_events;
  P->budget32 = (chunk_events == 0ull || chunk_events > 0xfffffffeull) ? 0xffffffffu : (uint32_t)chunk_events;
  P->end_eps = P->spec.end_time + 1e-9;
  for (int i = 0; i < P->spec.n_ing; ++i)
    for (int d = 0; d < P->spec.n_dc; ++d)
      for (int jt = 0; jt < 2; ++jt) {
        const double v = P->spec.transfer_s[i][d][jt];
        if (v == v && v < 1e300 && v > P->max_transfer) P->max_transfer = v;
      }
  P->max_transfer += dcsim_test_time_quantum; /* test hook: a rounded-up xfer_done instant may exceed t + transfer_s */
  P->state = (char*)calloc(n_replicas, (size_t)P->L.total_bytes);
  P->queues = (char*)calloc(n_replicas, (size_t)P->L.queue_bytes + 16);
  P->summary = out_summaries;
  P->lat_hist = lat_hist;
  { /* the arrival pre-pass and the list merge, one replica after the other */
    const size_t ne = n_replicas * (size_t)P->cap_arr;
    P->arr_t = (double*)calloc(ne, sizeof(double));
    P->arr_raw = (double*)calloc(ne, sizeof(double));
    P->arr_meta = (uint32_t*)calloc(ne, sizeof(uint32_t));
    P->arr_pred = (uint32_t*)calloc(ne, sizeof(uint32_t));
    P->arr_tx = (double*)calloc(ne, sizeof(double));
    P->arr_fin = (uint32_t*)calloc(ne, sizeof(uint32_t));
    P->ml_t = (double*)calloc(2 * ne, sizeof(double));
    P->ml_aux = (double*)calloc(2 * ne, sizeof(double));
    P->ml_meta = (uint32_t*)malloc(2 * ne * sizeof(uint32_t));
    memset(P->ml_meta, 0xee, 2 * ne * sizeof(uint32_t)); /* a position the merge never wrote would show */
    P->arr_hdr = (dcsim_arrhdr_t*)calloc(n_replicas, sizeof(dcsim_arrhdr_t));
    double clocks[2 * DCSIM_MAX_ING];
    uint32_t last[2 * DCSIM_MAX_ING];
    uint32_t ring[DCSIM_TRNG_RING];
    static dcsim_merge_ring_t merge_ring;
    if (rng_kind == 1) P->mt_state = (uint32_t*)calloc(n_replicas * (size_t)DCSIM_MT_N, sizeof(uint32_t));
    for (uint64_t r = 0; r < n_replicas; ++r) {
      if (rng_kind == 1) dcsim_generate_arrivals<true>(P, r, clocks, last, ring, 1); else dcsim_generate_arrivals<false>(P, r, clocks, last, ring, 1);
      dcsim_merge_arrivals(P, r, 0, &merge_ring);
      /* self-check: the merged list is a gap-free, (t)-sorted permutation of the replica's events */
      const dcsim_arrhdr_t* ah = P->arr_hdr + r;
      const double* mt = P->ml_t + 2 * r * (uint64_t)P->cap_arr;
      const uint32_t* mm = P->ml_meta + 2 * r * (uint64_t)P->cap_arr;
      for (uint32_t p = 0; p < ah->ml_count; ++p)
        if (mm[p] == 0xeeeeeeeeu || (p && mt[p] < mt[p - 1])) { fprintf(stderr, "hostemu: merged list broken at %u of %u (replica %llu)\n", p, ah->ml_count, (unsigned long long)r); abort(); }
    }
  }
  char* work = (char*)malloc((size_t)P->L.total_bytes);
  /* DCSIM_RECORDS=global: the kernel's "head staged" mode — only [0, rec_off) round-trips through the working copy,
   * the running-job records are used where they live (the block's home) */
  const char* rm = getenv("DCSIM_RECORDS");
  const bool head_only = rm && rm[0] == 'g';
  const size_t staged = head_only ? (size_t)P->L.rec_off : (size_t)P->L.total_bytes;
  long long total = 0;
  for (uint64_t r = 0; r < n_replicas; ++r) {
    char* home = P->state + r * (uint64_t)P->L.total_bytes;
    char* rec = head_only ? home : work;
    for (int guard = 0; guard < 100000000; ++guard) {
      const bool fresh = ((dcsim_hdr_t*)home)->initialized == 0u;
      if (!fresh) memcpy(work, home, staged); /* stage in */
      total += P->L.cap_stale ? (head_only ? dcsim_replica_step<true, true>(P, r, work, rec, fresh) : dcsim_replica_step<true, false>(P, r, work, rec, fresh))
                               : (head_only ? dcsim_replica_step<false, true>(P, r, work, rec, fresh) : dcsim_replica_step<false, false>(P, r, work, rec, fresh));
      memcpy(home, work, staged);             /* stage out */
      const dcsim_hdr_t* H = (const dcsim_hdr_t*)home;
      if (H->done || H->status || chunk_events == 0) break;
    }
  }
  free(work); free(P->state); free(P->queues); free(P->arr_t); free(P->arr_raw); free(P->arr_meta); free(P->arr_pred); free(P->arr_tx); free(P->arr_fin);
  free(P->ml_t); free(P->ml_aux); free(P->ml_meta); free(P->arr_hdr); free(P->mt_state); free(P);
  return total;
}

} /* extern "C" */
