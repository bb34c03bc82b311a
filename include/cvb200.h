/* cvb200.h -- C ABI of libcvb200.so: the B200-native Clairvoyante v3 / v3_slim
 * network (forward, loss, train step) for (33,4,4) pileup tensors.
 *
 * The reference has no FFI: its seam is the duck-typed Python object
 * `Clairvoyante` (clairvoyante/clairvoyante_v3.py:5-283,
 * clairvoyante_v3_slim.py:5-260) whose methods call TensorFlow's
 * `session.run`.  Every entry point below replaces one of those
 * `session.run` call sites; clairvoyante_b200/clairvoyante_v3.py binds them
 * with ctypes.  Plain pointers and sizes only -- no
 * torch / CUDA types in any signature (streams are passed as void*).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message
 *     is available from cvb_last_error() (thread-local).
 *   - x is float32 NHWC (n,33,4,4): element (i,h,w,c) at i*528+h*16+w*4+c
 *     (dataPrepScripts/CreateTensor.py:20-21,37-40; utils_v2.py:45), already
 *     channel-subtracted (utils_v2.py:46).
 *   - out16 is float32 (n,16): [base 4 | zygosity 2 | varType 4 | indelLength 6]
 *     = sigmoid / softmax outputs in the label order of callVar.py:54-57.
 *   - logits16 (optional, may be NULL) is the same layout before the final
 *     sigmoid / softmax: base pre-sigmoid, the others SELU(FC)+1e-10
 *     (clairvoyante_v3.py:124-137).
 *   - n == 0 is legal and a no-op (utils_v2.py:56-59 can yield empty batches).
 *   - a handle is bound to one CUDA device; calls may come from any host
 *     thread but only one call per handle may be in flight
 *     (callVar.py:197-205 contract).
 */
#ifndef CVB200_H
#define CVB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cvb_model cvb_model;

enum { CVB_V3 = 0, CVB_V3_SLIM = 1 };
/* arithmetic mode of the forward pass */
enum {
  CVB_COMPUTE_FP32 = 0,   /* fp32 SIMT everywhere (bit-faithful op order per layer) */
  CVB_COMPUTE_FP16X3 = 1, /* conv2/conv3/FC4/FC5+heads on tcgen05 with split-fp16 (hi+lo) operands, fp32 accumulate:
                             fp32-equivalent (logits within 1e-3 of fp64), the default */
  CVB_COMPUTE_FP16 = 2    /* v3_slim only (BASELINE configs[2] "v3_slim inference, fp16"): the same tensor pipeline with plain
                             fp16 operands -- one MMA term instead of three, activations kept as one fp16 plane, fp32
                             accumulate; logits within 2e-3 * max(1, max |logit|) of fp64 (tests/test_forward_gpu.py) */
};

const char* cvb_last_error(void);
int cvb_version(void);

/* replaces Clairvoyante.__init__/_buildGraph (clairvoyante_v3.py:7-175): allocates
 * parameters, Adam slots and work buffers on CUDA device `device`.              */
int cvb_create(int variant, int device, cvb_model** out);
/* replaces close()/__del__ (clairvoyante_v3.py:180,282) */
int cvb_destroy(cvb_model* m);

/* the 18 trainable variables by TF name (visualization.ipynb:100-121) */
int cvb_num_variables(const cvb_model* m);
int cvb_variable_info(const cvb_model* m, int idx, char* name, int name_cap,
                      int64_t* numel, int* ndim, int64_t dims[4]);
int64_t cvb_num_parameters(const cvb_model* m);
/* replaces restoreParameters / init (clairvoyante_v3.py:177,248): upload one variable.
 * slot 0 = value, 1 = Adam m, 2 = Adam v                                            */
int cvb_set_variable(cvb_model* m, const char* name, int slot, const float* host, int64_t n);
/* replaces saveParameters (clairvoyante_v3.py:243): download one variable */
int cvb_get_variable(cvb_model* m, const char* name, int slot, float* host, int64_t n);
/* replaces init() = session.run(tf.global_variables_initializer()) (clairvoyante_v3.py:177-178): draws every variable
 * from the reference's initialisers -- variance-scaling truncated normal (factor 2, FAN_IN) for conv1-3 / fc4 / fc5
 * kernels (:57,72,87,106,116), glorot-uniform for the four head kernels (:125-135), zero biases -- with a seeded
 * counter-based generator, zeroes the Adam slots and the step counter.  The reference is unseeded. */
int cvb_init_weights(cvb_model* m, uint64_t seed);
/* Adam step counter t (TF keeps beta1_power/beta2_power; t is their exponent) */
int cvb_set_step(cvb_model* m, int64_t t);
int cvb_get_step(const cvb_model* m, int64_t* t);
int cvb_set_compute_mode(cvb_model* m, int mode);

/* replaces predict/predictNoRT (clairvoyante_v3.py:257-280) with HOST buffers:
 * stages x through pinned memory, copies host->device, runs the kernels, copies
 * the results back into the four per-head arrays the reference's predict() returns.  x may be pageable or pinned.  Blocks until results are
 * in out16.                                                                      */
int cvb_predict_host(cvb_model* m, const float* x, int64_t n, float* base /*n,4*/, float* zygosity /*n,2*/,
                     float* var_type /*n,4*/, float* indel_length /*n,6*/, float* logits16 /*n,16 or NULL*/);
/* same call for candidate tensors the caller already holds as IEEE fp16 (`np.float16`, (n,33,4,4) C-contiguous): halves the
 * host->device bytes, which is what bounds cvb_predict_host on a PCIe-attached B200.  Exact whenever the values are
 * integers with |x| <= 2048 -- CreateTensor's counts are capped at 250 (dataPrepScripts/CreateTensor.py:296) -- because the
 * first device step widens them to the fp32 layout of cvb_predict_host; results are then bit-identical to it. */
int cvb_predict_host_f16(cvb_model* m, const uint16_t* x, int64_t n, float* base, float* zygosity, float* var_type,
                         float* indel_length, float* logits16);
/* the narrow feed (utils_v2.py:46 fused into the first device kernel): RAW CreateTensor counts -- the
 * non-negative integers of dataPrepScripts/CreateTensor.py:23-52 (`alnCode`), (n,33,4,4) C-contiguous, channel 0 NOT yet
 * subtracted -- as int16 (a half of the fp32 bytes on the host->device link) or uint8 (a quarter; the caller guarantees
 * every count <= 255, e.g. utils_v2.pack_counts checks it).  The first kernel widens each (row, base) position to fp32
 * and makes channels 1..3 relative to channel 0, exactly what GetTensor does on the host for the fp32 entry point
 * (integers: no rounding anywhere), so results are bit-identical to cvb_predict_host on the subtracted fp32 tensors. */
int cvb_predict_host_counts_i16(cvb_model* m, const int16_t* counts, int64_t n, float* base, float* zygosity,
                                float* var_type, float* indel_length, float* logits16);
int cvb_predict_host_counts_u8(cvb_model* m, const uint8_t* counts, int64_t n, float* base, float* zygosity,
                               float* var_type, float* indel_length, float* logits16);
/* element kinds of a candidate-tensor buffer */
enum { CVB_X_F32 = 0, CVB_X_F16 = 1, CVB_X_I16_COUNTS = 2, CVB_X_U8_COUNTS = 3 };
/* The same call split in two, for drivers that keep several small batches in flight (the reference runs predictNoRT on a
 * thread while it writes the previous batch, callVar.py:197-205): cvb_predict_submit copies x (n <= one device pass:
 * 18,944 sites for v3, 33,152 for v3_slim; any CVB_X_* element kind) into a staging slot, enqueues the upload, the kernels
 * and the download, and returns a ticket -- x may be reused at once; cvb_predict_collect waits for that ticket and fills
 * the four arrays (and logits16 if the ticket was submitted with want_logits).  Up to 4 tickets may be outstanding; they
 * complete in submission order; results are bit-identical to cvb_predict_host.  While a ticket is outstanding the
 * synchronous predict / train / set_variable entry points return an error instead of racing with it.               */
int cvb_predict_submit(cvb_model* m, const void* x, int x_kind, int64_t n, int want_logits, int* ticket);
int cvb_predict_collect(cvb_model* m, int ticket, float* base, float* zygosity, float* var_type, float* indel_length,
                        float* logits16);
/* same computation on DEVICE buffers (x, out16, logits16 are device pointers on the
 * handle's device); enqueued on `stream` (a cudaStream_t; NULL = the CUDA legacy default
 * stream, exactly as in the runtime API) and NOT synchronised.                                                  */
int cvb_predict_device(cvb_model* m, const float* x, int64_t n, float* out16, float* logits16,
                       void* stream);
/* cvb_predict_device for a device buffer of any element kind (CVB_X_*; x 16-byte aligned) */
int cvb_predict_device_x(cvb_model* m, const void* x, int x_kind, int64_t n, float* out16, float* logits16, void* stream);

/* replaces getLoss/getLossNoRT (clairvoyante_v3.py:207-227): forward with phase=False,
 * lambda=0; returns the SUM-over-batch loss (clairvoyante_v3.py:140-152).
 * x (n,33,4,4) and y (n,16) float32 host buffers.                                */
int cvb_loss_host(cvb_model* m, const float* x, const float* y, int64_t n, float* loss);

/* replaces train/trainNoRT (clairvoyante_v3.py:183-205): forward (SELU dropout on FC4
 * with `drop4`, selu.py:34-69), loss + l2*sum(0.5*||kernel||^2), backward, TF-1.x Adam
 * (clairvoyante_v3.py:174).  loss5 receives [loss, loss1..4-sum parts are folded: total,
 * base, zygosity, varType, indelLength] BEFORE the update, like session.run's fetch.
 * If apply_update == 0 the gradients are left in the gradient buffer (for a
 * data-parallel all-reduce by the caller) and cvb_apply_adam finishes the step.    */
int cvb_train_step_host(cvb_model* m, const float* x, const float* y, int64_t n,
                        float lr, float l2, float drop4, uint64_t dropout_seed,
                        int apply_update, float* loss6);
/* dropoutRateFC5 (clairvoyante_v3.py:121, selu.py:34-69, param.py:22; default 0): SELU dropout on FC5's output, which
 * feeds the zygosity / varType / indelLength heads, in every following cvb_train_step_host.  Its keep mask comes from the
 * same counter-based stream as FC4's (train_simt.cuh hash_uniform; host twin dropout_rng.py) under
 * dropout_seed ^ 0x5D5D5D5D5D5D5D5D, element index = site * N5 + unit. */
int cvb_set_dropout_fc5(cvb_model* m, float rate);
/* the same two calls for a batch held as fp16 values or as RAW int16 / uint8 counts (CVB_X_*; see cvb_predict_host_counts_*):
 * a half / a quarter of the bytes cross the host->device link and every micro-chunk is widened -- raw counts get
 * x[...,1:4] -= x[...,0:1], utils_v2.py:46 -- on the device; losses and gradients are bit-identical to the float32 call */
int cvb_loss_host_x(cvb_model* m, const void* x, int x_kind, const float* y, int64_t n, float* loss);
int cvb_train_step_host_x(cvb_model* m, const void* x, int x_kind, const float* y, int64_t n, float lr, float l2, float drop4,
                          uint64_t dropout_seed, int apply_update, float* loss6);
/* arithmetic of the large contractions of cvb_train_step_host / cvb_loss_host (train.py's model.train / getLoss path,
 * clairvoyante_v3.py:174,183-227).  FP32: fp32 SIMT kernels throughout.  BF16X3 (default): FC4 forward, data gradient and
 * weight gradient on tcgen05 with split-bf16 operands (x = hi + lo, three products, fp32 accumulate: ~2^-16 relative per
 * operand, fp32's exponent range).  BF16: the same kernels with the hi planes only -- BASELINE config "bf16 compute,
 * fp32 master weights".  Master weights, Adam slots, gradients and every other layer stay fp32 in all modes.          */
enum { CVB_TRAIN_FP32 = 0, CVB_TRAIN_BF16X3 = 1, CVB_TRAIN_BF16 = 2 };
int cvb_set_train_mode(cvb_model* m, int mode);
/* device pointer + element count of the flat fp32 gradient buffer (all 18 variables in
 * cvb_variable_info order, followed by 5 loss terms) for an external all-reduce     */
int cvb_grad_buffer(cvb_model* m, void** dev_ptr, int64_t* numel);
/* ---- data-parallel training (train.py over several GPUs): one process per GPU, every rank steps on its shard
 * of the global batch, and because the loss is a SUM over the batch (clairvoyante_v3.py:140-151) one SUM all-reduce of
 * [gradients | 4 loss sums] gives every rank the global step.  With a communicator attached the library issues that
 * ncclAllReduce itself, on its compute stream, inside cvb_train_step_host(apply_update = 1) between the backward pass and
 * the optimiser kernel -- no host synchronisation, nothing on a foreign stream.  NCCL is bound at run time (dlopen of
 * libnccl.so.2, or $CVB_NCCL_LIB): libcvb200.so does not link against it.
 *   cvb_nccl_unique_id : rank 0 draws the 128-byte ncclUniqueId and hands it to the other ranks by any means
 *                        (torch.distributed broadcast, MPI, a file)
 *   cvb_allreduce_init : every rank: ncclCommInitRank on the handle's device; the communicator belongs to the handle
 *   cvb_allreduce_attach : use an EXISTING ncclComm_t (owned by the caller; NULL detaches)
 *   cvb_allreduce_gradients : the reduction alone, for callers that drive apply_update = 0 / cvb_apply_adam themselves */
int cvb_nccl_unique_id(void* id128);
int cvb_allreduce_init(cvb_model* m, const void* id128, int nranks, int rank);
int cvb_allreduce_attach(cvb_model* m, void* nccl_comm);
int cvb_allreduce_gradients(cvb_model* m);
/* download gradients of one variable (testing) */
int cvb_get_gradient(cvb_model* m, const char* name, float* host, int64_t n);
/* finishes a step: loss6 (may be NULL) = [total, base, zygosity, varType, indelLength, lossL2] computed from the
 * (possibly all-reduced) loss sums and the PRE-update weights, then the TF-1.x Adam update on every variable */
int cvb_apply_adam(cvb_model* m, float lr, float l2, float* loss6);

/* ---- batch feed: native parser of the candidate-tensor text stream -------------------------------------------------
 * Replaces the per-row split()/np.array()/channel-subtract of utils_v2.GetTensor (clairvoyante/utils_v2.py:29-47) for
 * rows `chrom pos refseq v0 .. v527` (dataPrepScripts/CreateTensor.py:56).  Host code; needs no GPU.
 *
 * Parses at most max_lines COMPLETE lines from buf[0,len).  A trailing line without '\n' is parsed only if
 * final_chunk != 0, otherwise it is left for the next call (*consumed tells where to resume).
 * Per line i (0 <= i < *lines) meta[i*10 ..] = {status, line_off, line_len, chrom_off, chrom_len, pos_off, pos_len,
 * seq_off, seq_len, 0} (byte offsets into buf).  status: KEPT = 531 fields, all 528 values numeric, centre base
 * refseq[16] in ACGT (utils_v2.py:39); SKIPPED = well-formed but centre base not ACGT; MALFORMED = wrong field count
 * or a non-numeric value (the reference prints "UnpackATensorRecord Failure"); BLANK = whitespace only.
 * The KEPT rows are written compacted to x[0 .. *kept) as 528 floats each, already with channels 1..3 made relative
 * to channel 0 (utils_v2.py:46).  x must hold max_lines*528 floats, meta max_lines*10 int64.
 * threads <= 0: use the hardware concurrency (capped at 64 and at one thread per 64 lines).                         */
enum { CVB_LINE_KEPT = 0, CVB_LINE_SKIPPED = 1, CVB_LINE_MALFORMED = 2, CVB_LINE_BLANK = 3 };
int cvb_parse_tensor_text(const char* buf, int64_t len, int final_chunk, int64_t max_lines, int threads, float* x,
                          int64_t* meta, int64_t* lines, int64_t* kept, int64_t* consumed);

/* the position strings utils_v2.GetTensor yields (utils_v2.py:38-42), for the KEPT lines of a cvb_parse_tensor_text result
 * (same buf, its meta, its *lines): "chrom:pos:SEQ\n" each, SEQ upper-cased.  Returns the bytes written to out[0, cap) or -1
 * (cap = the consumed byte count of that parse is always enough). */
int64_t cvb_tensor_text_positions(const char* buf, const int64_t* meta, int64_t lines, char* out, int64_t cap);

/* fp32 candidate tensors -> the raw counts behind them for the narrow feed (cvb_predict_host_counts_*).  x = n_pos
 * positions of 4 channels ((n,33,4,4) has 132 n positions); subtracted != 0: channels 1..3 are relative to channel 0 as
 * GetTensor yields them (utils_v2.py:46), raw = x_i + x_0.  out_i16 always, out_u8 (may be NULL) saturating at 255.
 * *exact = 1 iff every count is an integer in [0, 32767] (otherwise keep the fp32 feed); *max_count <= 255 means the uint8
 * buffer is exact too.  Host code, threads <= 0: hardware concurrency. */
int cvb_pack_counts(const float* x, int64_t n_pos, int subtracted, int threads, int16_t* out_i16, uint8_t* out_u8,
                    int* max_count, int* exact);

/* CRC-32C (Castagnoli) of data[0,n) continuing from `crc` (0 to start): the checksum TensorFlow's checkpoint bundles
 * carry per tensor and per index block (tf.train.Saver, clairvoyante_v3.py:243-251).  Host code. */
uint32_t cvb_crc32c(uint32_t crc, const void* data, int64_t n);

/* Blosc-1 frames with LZ4 / LZ4HC blocks: what `blosc.pack_array(a, cname='lz4hc')` produced for the reference's
 * training set (clairvoyante/utils_v2.py:174-176) and `blosc.unpack_array` reads (:198,202).  Host code.
 * cvb_blosc_info: uncompressed / compressed size, typesize and flag byte of a frame (any pointer may be NULL).
 * cvb_blosc_decompress: frame -> dst[0, *out_n) (the pickled ndarray); byte-shuffle undone; other codecs are refused. */
int cvb_blosc_info(const void* frame, int64_t n, int64_t* nbytes, int64_t* cbytes, int* typesize, int* flags);
int cvb_blosc_decompress(const void* frame, int64_t n, void* dst, int64_t cap, int64_t* out_n);
/* the inverse, for `blosc.pack_array` (utils_v2.py:174-176,182-184): src[0, nbytes) -> one Blosc-1 frame with LZ4 streams
 * (byte-shuffled by typesize when do_shuffle != 0; stored uncompressed when that is not smaller).  cap >= cvb_blosc_compress_bound. */
int64_t cvb_blosc_compress_bound(int64_t nbytes);
int cvb_blosc_compress(const void* src, int64_t nbytes, int typesize, int do_shuffle, void* dst, int64_t cap, int64_t* out_n);

/* ---- alignment pile-up: SAM records -> candidate tensors ---------------------------------------------------------------
 * Replaces dataPrepScripts/CreateTensor.py (GenerateTensor :23-59, OutputAlnTensor :96-258), the producer of the tensor
 * text stream in the reference's pipelines (callVarBam.py:61).  Host code; needs no GPU.  One handle = one contig / region.
 *   ref_seq / ref_len : reference bases of the fetched region (what `samtools faidx` returned, header and newlines removed)
 *   ref_start         : 1-based position of ref_seq[0] (the reference's args.refStart), 0 = sequence starts at position 1
 *   cand_pos          : candidate positions (column 2 of the candidate list, 1-based), any order
 *   min_mq, dcov, min_coverage, consider_left_edge : the command-line options of CreateTensor.py:281-300
 * cvb_pileup_feed takes SAM text (`samtools view -F 2308` output, position-sorted) in arbitrary chunks; final_chunk != 0
 * flushes the centres still open.  Finished tensors queue up in ascending position order: cvb_pileup_take copies up to
 * max_sites of them as float32 (n,33,4,4) RAW counts (CreateTensor's alnCode: channel 0 is not yet subtracted,
 * utils_v2.py:46 does that) plus their centre positions.  stats = {SAM rows, rows used, malformed rows, open centres}.     */
typedef struct cvb_pileup cvb_pileup;
int cvb_pileup_create(const char* ref_seq, int64_t ref_len, int64_t ref_start, const int64_t* cand_pos, int64_t n_cand,
                      int min_mq, int dcov, int min_coverage, int consider_left_edge, cvb_pileup** out);
int cvb_pileup_destroy(cvb_pileup* p);
/* host threads for the CIGAR walks of one cvb_pileup_feed call (default 1): each thread owns a contiguous range of the
 * candidate list; results (tensors, their order, the statistics) are identical for every thread count */
int cvb_pileup_set_threads(cvb_pileup* p, int threads);
int cvb_pileup_feed(cvb_pileup* p, const char* sam, int64_t len, int final_chunk);
int64_t cvb_pileup_ready(const cvb_pileup* p);
int cvb_pileup_take(cvb_pileup* p, int64_t max_sites, float* x, int64_t* center, int64_t* n_out);
int cvb_pileup_stats(const cvb_pileup* p, int64_t stats[4]);
/* the text rows of CreateTensor.py:56 for n tensors ("ctg pos refseq33 v0 .. v527\n", values "%0.1f"); returns the number of
 * bytes written to out[0, cap) or -1 (cap must allow 8.6 KB + strlen(ctg) per row) */
int64_t cvb_pileup_format_rows(const char* ctg, const int64_t* center, const float* x, int64_t n, const char* ref_seq,
                               int64_t ref_len, int64_t ref_start, char* out, int64_t cap);

/* ---- variant-candidate extraction: SAM records -> candidate positions ------------------------------------------------
 * Replaces dataPrepScripts/ExtractVariantCandidates.py (OutputCandidate :22-42, MakeCandidates :53-253), the first stage
 * of the reference's calling pipeline (callVarBam.py:56-66).  Host code; needs no GPU.  One handle = one contig / region.
 *   ctg_name                 : rows whose RNAME differs are skipped (:133-135)
 *   ref_seq, ref_len, ref_start : as for cvb_pileup_create
 *   ctg_start, ctg_end       : the region test of :187-188 exactly as the reference evaluates it -- 0-based position p is
 *                              kept when ctg_start <= p <= ctg_end, where ctg_start is the --ctgStart argument + 1 (:63);
 *                              pass -1, -1 for no region
 *   bed_begin/bed_end/n_bed  : half-open [begin, end) intervals of this contig as the reference builds them (:95-98:
 *                              end = column 3 - 1, +1 if that equals begin); n_bed = -1 for no BED file
 *   min_mq, min_coverage, threshold : --minMQ, --minCoverage, --threshold
 *   output_prob, seed        : --gen4Training subsample probability (< 0: keep everything), hash seed
 * cvb_candidates_feed takes position-sorted SAM text in arbitrary chunks (final_chunk != 0 ends the input);
 * cvb_candidates_take moves out the finished rows ("ctg pos refBase total k0 n0 .. k6 n6\n") and their 1-based positions.
 * stats = {SAM rows, reads processed, malformed rows, open positions}.                                                  */
typedef struct cvb_candidates cvb_candidates;
int cvb_candidates_create(const char* ctg_name, const char* ref_seq, int64_t ref_len, int64_t ref_start, int64_t ctg_start,
                          int64_t ctg_end, const int64_t* bed_begin, const int64_t* bed_end, int64_t n_bed, int min_mq,
                          double min_coverage, double threshold, double output_prob, uint64_t seed, cvb_candidates** out);
int cvb_candidates_destroy(cvb_candidates* c);
/* host threads per cvb_candidates_feed call (default 1): rows are tokenised in parallel and the counting is split by reference
 * position; rows, their order and the statistics are identical for every thread count */
int cvb_candidates_set_threads(cvb_candidates* c, int threads);
int cvb_candidates_feed(cvb_candidates* c, const char* sam, int64_t len, int final_chunk);
int64_t cvb_candidates_pending_bytes(const cvb_candidates* c);
int64_t cvb_candidates_pending(const cvb_candidates* c);
int cvb_candidates_take(cvb_candidates* c, char* text, int64_t text_cap, int64_t* text_len, int64_t* pos, int64_t pos_cap,
                        int64_t* n_pos);
int cvb_candidates_stats(const cvb_candidates* c, int64_t stats[4]);

/* ---- VCF records of one batch of calls: the per-site loop of `Output` (clairvoyante/callVar.py:58-153) --------------------
 * x (n,33,4,4) the batch as GetTensor yields it; pos = its n position strings "chrom:pos:SEQ" separated by '\n';
 * base (n,4), z (n,2), t (n,4), l (n,6) the four head outputs; show_ref as --showRef; qual_cut = --qual or -1.
 * Writes the records (one line each, sites without a record skipped) to out[0, cap) and returns the byte count, -1 on error
 * (cap >= pos_len + 256 * n is always enough).  Host code. */
int64_t cvb_vcf_records(const float* x, const char* pos, int64_t pos_len, const float* base, const float* z, const float* t,
                        const float* l, int64_t n, int show_ref, int qual_cut, char* out, int64_t cap);

/* `samtools view -F <flag_mask> <file> ctg[:start-end]` for SAM TEXT (CreateTensor.py:134-136, ExtractVariantCandidates.py:
 * 107-109 run it with -F 2308): copies the complete lines of in[0, len) that samtools would print to out (capacity len + 1)
 * -- header lines, other contigs, records with flag & flag_mask and records that do not overlap the 1-based inclusive region
 * [start, end] are dropped; start < 0 = no region.  A trailing line without '\n' is processed only if final_chunk != 0;
 * *consumed = bytes of `in` that were processed (feed the rest again with the next chunk).  Host code. */
int cvb_sam_view(const char* in, int64_t len, int final_chunk, const char* ctg_name, int flag_mask, int64_t start, int64_t end,
                 char* out, int64_t* out_len, int64_t* consumed);

/* pinned host memory helpers for the batch feed (utils_v2.GetTensor replacement) */
int cvb_alloc_pinned(int64_t bytes, void** out);
int cvb_free_pinned(void* p);

/* test aid: copy the first n floats of an intermediate of the LAST device pass to host.
 * which: 0 = p2 (pooled conv2, padded rows), 1 = p3 (pooled conv3 = FC4 input), 2 = h4 (FC4 output) */
int cvb_debug_read(cvb_model* m, int which, float* host, int64_t n);

/* per-kernel device timing for bench.py's roofline: while enabled, every kernel launch of
 * the forward pass is bracketed by CUDA events on the launching stream.  cvb_profile_read
 * synchronises the device and returns, per kernel kind,
 * the summed duration in ms and the number of launches since cvb_profile_begin.
 * kinds: 0 SIMT front (conv1[+conv2]), 1 tcgen05 conv2, 2 conv3, 3 fc4, 4 tail (FC5 + heads)        */
int cvb_profile_begin(cvb_model* m);
int cvb_profile_read(cvb_model* m, double ms[5], int64_t launches[5]);

/* counters for bench.py: number of this library's kernels launched so far on the handle */
int64_t cvb_kernel_launches(const cvb_model* m);

#ifdef __cplusplus
}
#endif
#endif
