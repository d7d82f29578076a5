/*
 * lwb200.h -- C ABI of liblwb200.so, the B200 (sm_100a) implementation of the Lightworks
 * emulator probability hot path.
 *
 * The reference (Aegiq/lightworks v2.3.3) is pure Python and has no FFI; the interface each
 * entry point replaces is therefore a Python method, cited per function as file:line relative
 * to the reference root.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative LWB_E_* code otherwise;
 *     lwb_last_error() gives the message for the calling thread.  No C++ exception crosses the ABI.
 *   - complex numbers are interleaved (re, im) float64 pairs ("complex128"), row-major.
 *   - U[j*m + i] is the amplitude input mode i -> output mode j (rows = outputs), exactly the
 *     reference convention (permanent.py:175-178, slos.py:100).
 *   - Fock states are uint8 occupation vectors of length m.
 *   - `dtype` selects the ARITHMETIC precision: LWB_C128 (complex128) or LWB_C64 (complex64,
 *     inputs are rounded to float on the device, results are returned as float64).
 *   - entry points without a suffix take HOST pointers, copy in/out and return when the result is
 *     in the caller's buffer.  `_dev` entry points take DEVICE pointers, enqueue on the handle's
 *     stream and return without synchronising (for callers that own device memory / streams).
 *   - the library never keeps a caller pointer after the call returns.
 */
#ifndef LWB200_H
#define LWB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lwb_context* lwb_handle;

enum { LWB_C128 = 0, LWB_C64 = 1 };
enum {
    LWB_OK = 0,
    LWB_E_ARG = -1,      /* bad argument (ValueError in the Python shim)             */
    LWB_E_PHOTONS = -2,  /* photon-number mismatch (PhotonNumberError / ValueError)  */
    LWB_E_LIMIT = -3,    /* size beyond what the kernels are built for              */
    LWB_E_CUDA = -4,     /* CUDA runtime error (BackendError)                       */
    LWB_E_NOMEM = -5,
    LWB_E_COMM = -6      /* NCCL missing or a collective failed (BackendError)        */
};

/* ---- lifecycle ---------------------------------------------------------------------------- */
int lwb_create(int device, lwb_handle* out);
int lwb_destroy(lwb_handle h);
const char* lwb_last_error(void);
int lwb_version(void);
/* Run on a caller-owned CUDA stream (cudaStream_t as void*; NULL = the legacy default stream).
 * use_own != 0 restores the handle's own non-blocking stream. */
int lwb_set_stream(lwb_handle h, void* cuda_stream, int use_own);
int lwb_synchronize(lwb_handle h);
/* Device facts for roofline reporting. */
int lwb_device_info(lwb_handle h, int* sm_count, int* sm_clock_khz, char* name, int name_len);
/* Kernel variant selection for n-photon batched permanents (sweeps): lanes per permanent = 2^log_l, from the
 * variants compiled into the library; a negative log_l restores the default. */
int lwb_set_tuning(int n, int log_l);
/* Process-wide options.  "multiplicity_aware" (default 1): full-basis probability ranges use the
 * multiplicity-aware Glynn walk (prod(s_i+1)/2 terms for an output with occupations s_i) instead of
 * expanding repeated rows into 2^(n-1) terms; 0 forces the plain expanded-row kernel.
 * "slos_units" (default 4): which kernel runs a SLOS layer (csrc/slos_units.cu).  4: record kernel - prefix units,
 * coalesced parent streams, parent blocks in shared memory, host-built records instead of index arithmetic, terms
 * accumulated with FMAs (rounding differs from the other two by ~1e-16 relative).  1: unit kernel - same units,
 * indices derived per child, rounds exactly like the gather kernel.  0: gather-form kernel (csrc/slos_kernels.cuh),
 * which rounds like the reference's Python complex arithmetic.
 * "source_string_keys" (default 0, test hook): lwb_source_statistics keys its classes by strings even when the
 * packed 62-bit form fits. */
int lwb_set_option(const char* name, int value);
/* Host-only view of the multiplicity-aware walk for n photons (2 <= n <= 16): pattern = index of a partition of
 * n (descending lexicographic order); returns its multiplicities (n_distinct entries, descending), the number of
 * terms of the walk, the symmetry factor (2 if halved) and the sum of |coefficients| (= 2^n / factor). */
int lwb_multiplicity_plan(int n, int pattern, int* n_patterns, int* multiplicities, int* n_distinct, int* n_terms,
                          int* symmetry_factor, double* abs_coef_sum);
/* Host-only self-check of the SLOS prefix-unit plan of layer (m modes, k photons): verifies every child / parent
 * index rule of csrc/slos_units.cu against exact lex ranks; *children = states checked (= C(m+k-1, k)). */
int lwb_slos_plan_check(int m, int k, uint64_t* units, uint64_t* items, uint64_t* children);
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int lwb_launch_count(lwb_handle h, uint64_t* count);

/* ---- Fock-state indexing (host, exact) ------------------------------------------------------
 * Replaces fock_basis/_sums, lightworks/emulator/utils/state.py:22-35 (ordering contract). */
int lwb_fock_count(int m, int n, uint64_t* count);
int lwb_fock_rank(int m, const uint8_t* state, uint64_t* rank);
int lwb_fock_unrank(int m, int n, uint64_t rank, uint8_t* state);
/* SLOS (dict insertion) order of slos.py:96-103 */
int lwb_slos_rank(int m, const uint8_t* state, uint64_t* rank);
int lwb_slos_unrank(int m, int n, uint64_t rank, uint8_t* state);
/* Occupation vectors of `count` states given by their positions in the SLOS order (host). */
int lwb_slos_states(int m, int n, const uint64_t* ranks, uint64_t count, uint8_t* states);
/* States [r0, r0+cnt) of fock_basis(m, n) enumerated on the device; states: cnt*m bytes. */
int lwb_fock_basis(lwb_handle h, int m, int n, uint64_t r0, uint64_t cnt, uint8_t* states);
int lwb_fock_basis_dev(lwb_handle h, int m, int n, uint64_t r0, uint64_t cnt, uint8_t* d_states);

/* ---- permanents -----------------------------------------------------------------------------
 * thewalrus.perm semantics (call site permanent.py:81): B dense n x n matrices -> B permanents. */
int lwb_perm_matrices(lwb_handle h, const double* A, uint64_t B, int n, int dtype, double* out);
int lwb_perm_matrices_dev(lwb_handle h, const double* d_A, uint64_t B, int n, int dtype, double* d_out);

/* PermanentBackend.probability_amplitude / probability for every (input, output) pair:
 * permanent.py:53-114, looped by simulator.py:98-104 and analyzer.py:120-151.
 *   ins  [n_in][m], outs [n_out][m]; every state must hold the same photon number.
 *   want_probs = 0: out is complex [n_in][n_out];  1: out is real |amp|^2 [n_in][n_out]. */
int lwb_perm_amplitudes(lwb_handle h, const double* U, int m, const uint8_t* ins, int n_in,
                        const uint8_t* outs, uint64_t n_out, int dtype, int want_probs, double* out);
int lwb_perm_amplitudes_dev(lwb_handle h, const double* d_U, int m, const uint8_t* d_ins, int n_in,
                            const uint8_t* d_outs, uint64_t n_out, int n_photons, int dtype,
                            int want_probs, double* d_out);

/* Batch across unitaries: lightworks.Batch (sdk/tasks/batch.py:45-58) holds tasks that typically differ only in the
 * circuit parameters; the reference runs them one by one (emulator/backends/backend.py:102-103).  n_tasks unitaries
 * U[n_tasks][m][m], inputs ins[n_tasks][n_in][m] (or one shared set [n_in][m] when shared_inputs != 0), shared outputs
 * outs[n_out][m], all with the same photon number; out[n_tasks][n_in][n_out] (complex or |amp|^2).  One launch. */
int lwb_perm_amplitudes_batch(lwb_handle h, const double* U, int n_tasks, int m, const uint8_t* ins, int n_in,
                              int shared_inputs, const uint8_t* outs, uint64_t n_out, int dtype, int want_probs, double* out);

/* The loop body of PermanentBackend.full_probability_distribution (permanent.py:145-150) over the
 * rank range [r0, r0+cnt) of fock_basis(m, n): probs[k] = |<basis[r0+k]|U|in>|^2.  Output states
 * are unranked on the device; rank ranges are the multi-GPU shard. */
int lwb_perm_basis_probs(lwb_handle h, const double* U, int m, const uint8_t* in_state, uint64_t r0,
                         uint64_t cnt, int dtype, double* probs);
int lwb_perm_basis_probs_dev(lwb_handle h, const double* d_U, int m, const uint8_t* d_in_state,
                             int n_photons, uint64_t r0, uint64_t cnt, int dtype, double* d_probs);

/* One large permanent, Gray-code space split over every SM.  [q_frac_begin, q_frac_end) selects a
 * slice rank/size of the Gray space for multi-GPU runs (0,1 = all); the caller sums the slices. */
int lwb_perm_single(lwb_handle h, const double* A, int n, int dtype, int slice, int n_slices, double* out2);
int lwb_perm_single_dev(lwb_handle h, const double* d_A, int n, int dtype, int slice, int n_slices,
                        double* d_out2);

/* ---- SLOS ------------------------------------------------------------------------------------
 * SLOSBackend.calculate (slos.py:82-105): amplitude of every output state of C(m+n-1, n).
 *   order 0: fock_basis (reference permanent) order; order 1: SLOS dict-insertion order.
 *   amps: S complex.  in_state is always a HOST pointer (it drives the layer loop). */
int lwb_slos_amplitudes(lwb_handle h, const double* U, int m, const uint8_t* in_state, int order, double* amps);
int lwb_slos_amplitudes_dev(lwb_handle h, const double* d_U, int m, const uint8_t* in_state, int order,
                            double* d_amps);
/* |amplitude|^2 of every output state, same orders (the loop body of slos.py:73-74 without filtering). */
int lwb_slos_basis_probs(lwb_handle h, const double* U, int m, const uint8_t* in_state, int order, double* probs);
int lwb_slos_basis_probs_dev(lwb_handle h, const double* d_U, int m, const uint8_t* in_state, int order,
                             double* d_probs);

/* ---- full distributions -------------------------------------------------------------------------
 * {Permanent,SLOS}Backend.full_probability_distribution (permanent.py:116-163, slos.py:43-80) for a
 * compiled circuit with m_tot = n_modes + loss_modes modes: vacuum shortcut, loss-mode zero padding,
 * zero-real-photon skip (permanent only), strict p > threshold, loss-mode marginalisation, 1 - sum
 * vacuum entry (permanent only).  Entries come back in the reference's dict insertion order:
 *   keys [count][n_modes] uint8, vals [count].  backend: 0 = permanent, 1 = slos.
 * Returns LWB_E_LIMIT if more than `cap` entries would be produced (count_out is still set). */
enum { LWB_BACKEND_PERMANENT = 0, LWB_BACKEND_SLOS = 1 };
int lwb_pdist(lwb_handle h, const double* U_full, int m_tot, int n_modes, const uint8_t* in_state,
              double threshold, int backend, int dtype, uint8_t* keys, double* vals, uint64_t cap,
              uint64_t* count_out);

/* ---- on-device sampling ---------------------------------------------------------------------------
 * SamplerRunner._gen_samples_from_dist (simulation/sampler.py:329-349): rng.choice(len(dist), p=probs, size=N),
 * i.e. numpy's  cdf = cumsum(p) / sum(p); index = searchsorted(cdf, rng.random(N), side='right').
 * The caller supplies the N uniform variates (numpy's Generator on the host keeps the reference's seeding);
 * the CDF is built and searched on the device.  Entries with p <= threshold count as zero, which is how the
 * reference's dict drops them (permanent.py:151, slos.py:74); pass a negative threshold to keep everything.
 *   index[k] = position in `probs` drawn by uniform[k];  total (optional) = sum of the kept probabilities.
 * lwb_basis_sample computes the full-basis distribution of one input state on the device (backend 0: permanent
 * path, indices in fock_basis order; 1: SLOS path, indices in SLOS dict order) and samples it without the
 * probabilities ever leaving the GPU: host traffic is m*m*16 + 8*N bytes in, 8*N + 8 bytes out. */
int lwb_sample(lwb_handle h, const double* probs, uint64_t S, double threshold, const double* uniform,
               uint64_t n_samples, uint64_t* index, double* total);
int lwb_sample_dev(lwb_handle h, const double* d_probs, uint64_t S, double threshold, const double* d_uniform,
                   uint64_t n_samples, uint64_t* d_index, double* d_total);
int lwb_basis_sample(lwb_handle h, int backend, const double* U, int m, const uint8_t* in_state, double threshold,
                     const double* uniform, uint64_t n_samples, uint64_t* index, double* total);
/* lwb_basis_sample restricted to the output states SamplerRunner._sample_N_outputs keeps (simulation/sampler.py:275-327):
 * optional threshold detection (s_i -> min(s_i, 1)), heralded modes holding exactly herald_counts photons, at least
 * min_detection photons outside the heralded modes, and post-selection rules (sdk/utils/post_selection.py:105-138)
 * "the photons in modes rule_modes[rule_mode_offsets[q] .. rule_mode_offsets[q+1]) number one of
 * rule_counts[rule_count_offsets[q] ..)" - modes are FULL-circuit mode indices.  A device kernel zeroes the excluded
 * states before the CDF is built; *total = probability mass of the kept states. */
int lwb_basis_sample_selected(lwb_handle h, int backend, const double* U, int m, const uint8_t* in_state, double threshold,
                              int threshold_detect, int min_detection, const int32_t* herald_modes,
                              const int32_t* herald_counts, int n_heralds, const int32_t* rule_modes,
                              const int32_t* rule_mode_offsets, const int32_t* rule_counts, const int32_t* rule_count_offsets,
                              int n_rules, const double* uniform, uint64_t n_samples, uint64_t* index, double* total);

/* ---- dense distributions and their merge-convolution -------------------------------------------------
 * pdist_calc / annotated_state_pdist_calc (simulation/probability_distribution.py:25-164): the output distribution
 * of an imperfect source is a weighted sum, over the source's annotated inputs, of the CONVOLUTION of the
 * distributions of the mutually distinguishable photon groups ("new_pdist[output.merge(result)] += p1 * p2",
 * :150-158).  A distribution lives on the device as a dense float64 vector over the key space
 *     K(n_modes, kmax) = occupation vectors of n_modes modes holding 0..kmax photons,
 *     index(s) = sum_{j<k} C(n_modes+j-1, j) + fock_basis rank of s among the k-photon states (k = photons of s),
 * addressed through an integer slot of the handle.
 *   lwb_dist_from_circuit: {Permanent,SLOS}Backend.full_probability_distribution (permanent.py:116-163,
 *       slos.py:43-80; threshold, loss-mode marginalisation, vacuum entry) straight into a slot (nothing returns
 *       to the host);
 *   lwb_dist_convolve: out = a (*) b, new slot over K(n_modes, kmax_a + kmax_b);
 *   lwb_dist_axpy: acc += alpha * x (probability_distribution.py:160-164 / :62-69);
 *   lwb_dist_axpy_many: acc += alpha_j * x_j for j = 0 .. n-1 in that order, one launch (the per-input accumulation
 *       loop of probability_distribution.py:137-164 for every annotated input that shares its largest photon group);
 *   lwb_dist_sample: lwb_sample on a slot;  lwb_dist_download: the dense vector (size from lwb_dist_info). */
int lwb_keyspace_size(int n_modes, int kmax, uint64_t* size);
int lwb_keyspace_index(int n_modes, const uint8_t* state, uint64_t* index);
int lwb_keyspace_state(int n_modes, int kmax, uint64_t index, uint8_t* state);
int lwb_keyspace_states(int n_modes, int kmax, const uint64_t* indices, uint64_t count, uint8_t* states);
int lwb_dist_create(lwb_handle h, int n_modes, int kmax, const double* dense_or_null, int* slot);
int lwb_dist_from_circuit(lwb_handle h, const double* U_full, int m_tot, int n_modes, const uint8_t* in_state,
                          double threshold, int backend, int dtype, int* slot);
int lwb_dist_convolve(lwb_handle h, int slot_a, int slot_b, int* slot_out);
int lwb_dist_axpy(lwb_handle h, int slot_acc, double alpha, int slot_x);
int lwb_dist_axpy_many(lwb_handle h, int slot_acc, int n_terms, const double* alphas, const int* slots);
int lwb_dist_info(lwb_handle h, int slot, int* n_modes, int* kmax, uint64_t* size, void** device_ptr);
int lwb_dist_download(lwb_handle h, int slot, double* dense);
int lwb_dist_sample(lwb_handle h, int slot, const double* uniform, uint64_t n_samples, uint64_t* index, double* total);
int lwb_dist_free(lwb_handle h, int slot);

/* ---- imperfect-source input statistics (host only) ------------------------------------------------
 * Source._build_statistics (emulator/components/source.py:134-161 with :163-406): every input the source can emit
 * for the target `state` and its probability, thresholded (p >= threshold) and renormalised, in the reference's
 * dict order.  Entry k is two occupation vectors: group_occ[k] = photons per mode that share label 0 (mutually
 * indistinguishable), single_occ[k] = photons per mode that carry a label of their own; *annotated = 0 when the
 * reference would return plain States (purity = indistinguishability = 1: group_occ is the state).  Returns
 * LWB_E_LIMIT with *count set when cap is too small. */
int lwb_source_statistics(int n_modes, const uint8_t* state, double purity, double brightness,
                          double indistinguishability, double threshold, uint64_t cap, uint8_t* group_occ,
                          uint8_t* single_occ, double* probs, uint64_t* count, int* annotated);

/* ---- multi-GPU (SURVEY.md section 8b "lwb_init(n_gpus)", 8e) --------------------------------------------
 * The path shards where the reference loops over independent units: output Fock states of one distribution
 * (permanent.py:145-157), the Gray-code space of one large permanent (thewalrus.perm, call site permanent.py:81),
 * and the unique sub-inputs of an imperfect source (probability_distribution.py:127-133).  A communicator is
 *   LOCAL: one process drives n GPUs - lwb_comm_init_local(n_gpus) creates one handle per device (this is SURVEY
 *          8b's lwb_init(n_gpus)); what lightworks_b200.install(n_gpus=N) puts behind Backend("permanent"); or
 *   RANK:  one process per GPU - lwb_comm_init_rank(handle, rank, world, id) with the 128-byte NCCL unique id from
 *          lwb_comm_unique_id() on rank 0, broadcast by the host framework (torchrun, MPI, a file).
 * NCCL is dlopen'ed ("libnccl.so.2"): no link-time dependency, shares the NCCL the host process already loaded.
 * Shards of a distribution are contiguous rank ranges of fock_basis(m, n) holding equal WORK of the
 * multiplicity-aware walk (lwb_shard_cuts).  When the GPUs can map each other's memory (NVLink peer access, CUDA IPC
 * between processes) the permanent kernels store every probability into the peers' copies as it is produced and the
 * step ends with the all-reduce of the sum, which doubles as the completion barrier; otherwise the shards travel by
 * in-place NCCL broadcasts.  Every rank (LOCAL: device 0) ends with the WHOLE distribution, contiguous, in
 * fock_basis order. */
typedef struct lwb_comm_s* lwb_comm;
int lwb_comm_unique_id(uint8_t* id128);
int lwb_comm_init_rank(lwb_handle h, int rank, int world, const uint8_t* id128, lwb_comm* out);
int lwb_comm_init_local(int n_gpus, lwb_comm* out);
int lwb_comm_destroy(lwb_comm c);
/* kind: 0 LOCAL, 1 RANK; peer_stores: 1 when kernels write into the peers' buffers directly */
int lwb_comm_info(lwb_comm c, int* rank, int* world, int* kind, int* peer_stores);
/* "peer_stores" 0/1: force the NCCL-broadcast (RANK) / peer-copy (LOCAL) exchange instead of in-kernel peer stores */
int lwb_comm_set_option(lwb_comm c, const char* name, int value);
/* LOCAL: the handle of device `index` (owned by the communicator); RANK: index 0 = the caller's handle */
int lwb_comm_handle(lwb_comm c, int index, lwb_handle* out);
/* Ranks of fock_basis(m, n) at which the cumulative work of the permanent walk reaches the given ascending
 * fractions in [0, 1] (host, exact integer arithmetic; n < 7 or n > 16: uniform cost per state). */
int lwb_shard_cuts(int m, int n, int n_fractions, const double* fractions, uint64_t* cuts);
int lwb_comm_shard(lwb_comm c, int m, int n, int rank, uint64_t* r0, uint64_t* r1);
/* lwb_perm_basis_probs over the whole basis, sharded.  probs may be NULL (result stays on the device).
 *   LOCAL: probs receives all S probabilities.  RANK: all S when gather_host != 0, else only this rank's shard,
 *   written at probs + r0 (every rank's DEVICE copy is complete either way).  sum_out: sum of all probabilities. */
int lwb_comm_perm_basis_probs(lwb_comm c, const double* U, int m, const uint8_t* in_state, int dtype, int gather_host,
                              double* probs, double* sum_out);
/* RANK only, device-resident inputs of this rank, asynchronous on the handle's stream: *d_probs = this rank's
 * complete copy (library-owned, S doubles), *d_sum = device scalar holding the all-reduced sum. */
int lwb_comm_perm_basis_probs_dev(lwb_comm c, const double* d_U, int m, const uint8_t* d_in_state, int n_photons,
                                  int dtype, double** d_probs, double** d_sum);
/* lwb_pdist(backend = permanent) with the permanents sharded over the communicator. */
int lwb_comm_pdist(lwb_comm c, const double* U_full, int m_tot, int n_modes, const uint8_t* in_state, double threshold,
                   int dtype, uint8_t* keys, double* vals, uint64_t cap, uint64_t* count_out);
/* lwb_perm_single with the Gray-code space split over the communicator and the 16-byte sum (NCCL all-reduce for
 * RANK, fixed-order host sum for LOCAL).
 * Below 24 photons one GPU is faster than the exchange: every rank (RANK) / device 0 (LOCAL) walks the whole space. */
int lwb_comm_perm_single(lwb_comm c, const double* A, int n, int dtype, double* out2);
int lwb_comm_perm_single_dev(lwb_comm c, const double* d_A, int n, int dtype, double* d_out2);
/* lwb_dist_from_circuit for `count` input states at once (probability_distribution.py:127-133: the unique
 * sub-inputs of an imperfect source).  h: the handle that receives the slots.  comm may be NULL (everything on h,
 * one synchronisation at the end) or a LOCAL communicator whose device 0 is h: input k is computed on device
 * k mod n_gpus and its dense distribution copied to h's device over NVLink. */
int lwb_dists_from_circuit(lwb_handle h, lwb_comm comm, const double* U_full, int m_tot, int n_modes,
                           const uint8_t* in_states, int count, double threshold, int backend, int dtype, int* slots);

/* Debugging aid: returns (and clears) the calling thread's pending CUDA runtime error code, message into msg. */
int lwb_debug_pending_cuda_error(char* msg, int len);

/* ---- micro-benchmarks used for the roofline denominators ------------------------------------ */
/* Dependent-free DFMA stream on every SM: returns achieved FP64 FLOP/s (2 flops per DFMA). */
int lwb_fp64_peak(lwb_handle h, double* flops_per_s);
/* FP64 issue-rate probes for register-operand patterns (1: 3-register DFMA, 2: DFMA with a shared
 * operand, 3: DMUL, 4: DADD, 5: complex-multiply chain); reported as instructions x 2 flops / s. */
int lwb_fp64_probe(lwb_handle h, int mode, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* LWB200_H */
