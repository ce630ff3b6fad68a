/*
 *  fiasco_b200.h -- C ABI of the B200-native FIASCO encoder hot path.
 *
 *  This is the drop-in boundary below the public libfiasco API: one call encodes the
 *  bintree subdivision of whole frames / independent tiles on the GPU.  It replaces, in
 *  the reference coder, the call
 *
 *      costs = subdivide (MAXCOSTS, GRAY, RANGE, &range, wfa, c, ..., NO);
 *                                              /root/reference/codec/coder.c:743 (grey)
 *                                              /root/reference/codec/coder.c:805 (per band)
 *
 *  together with everything that call reaches: codec/subdivide.c:60 (subdivide),
 *  codec/approx.c:74,317,644 (approximate_range, matching_pursuit, orthogonalize),
 *  codec/ip.c:72,184 (compute_ip_images_state, compute_ip_states_state),
 *  codec/domain-pool.c:707-852 (rle pool: generate/bits/update/append),
 *  codec/coeff.c:215-267 (adaptive coefficient model), codec/control.c:48 (append_state),
 *  codec/bintree.c:35,55 (tree model), lib/rpf.c:59,113 (rtob/btor).
 *
 *  Plain C types only: pointers and sizes, no CUDA or torch types in any signature
 *  (a cudaStream_t travels as void *).  All functions return 0 on success and a
 *  non-zero FB200_E* code on failure, with a message in the caller's err buffer --
 *  the host library maps that to libfiasco's "return 0 + fiasco_get_error_message()"
 *  convention (/root/reference/lib/error.c:113,178).
 *
 *  There is NO CPU fallback: if no CUDA device / driver is present every entry point
 *  that needs one fails with FB200_ENODEVICE.
 */
#ifndef FIASCO_B200_H
#define FIASCO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_MAXEDGES	5	/* codec/wfa.h:20 */
#define FB200_MAXSTATES 6000	/* codec/wfa.h:21 */
#define FB200_MAXLABELS 2	/* codec/wfa.h:22 */
#define FB200_MAXLEVEL	22	/* codec/wfa.h:23 */

enum
{
   FB200_OK	     = 0,
   FB200_EINVAL	     = 1,	/* bad parameter */
   FB200_ENODEVICE   = 2,	/* no usable CUDA device */
   FB200_ECUDA	     = 3,	/* CUDA runtime error (message has the details) */
   FB200_ECAPACITY   = 4,	/* a tile needed more states than 'state_capacity' */
   FB200_EMAXSTATES  = 5,	/* "Maximum number of states reached!" (control.c:129) */
   FB200_ENOROOT     = 6,	/* "No root state generated!" (coder.c:751) */
   FB200_EUNSUPPORTED = 7	/* option outside the supported set (see below) */
};

/*
 *  Coder parameters of one job; every tile of a job shares them.  The values are the
 *  ones found in the reference's coding_t / c_options_t AFTER alloc_coder()'s clamping
 *  (codec/coder.c:249-327); fb200_params_init() performs exactly that clamping from
 *  user-level options.
 */
typedef struct fb200_params
{
   int	 width, height;		/* tile geometry in pixels (even numbers) */
   int	 bands;			/* 1 = grey; 3 = Y,Cb,Cr 4:4:4 (lib/image.c:348) */
   int	 level;			/* bintree level of the tile (coder.c:249-256) */
   int	 lc_min_level;		/* smallest range level (c->options.lc_min_level) */
   int	 lc_max_level;		/* largest range level */
   int	 images_level;		/* state images kept up to this level (5) */
   int	 max_elements;		/* edges per linear combination, 1..5 */
   int	 max_states;		/* dictionary size (wfainfo->max_states) */
   int	 chroma_max_states;	/* chroma dictionary size (40) */
   float price;			/* 128 * 64 / quality (coder.c:164) */
   float chroma_decrease;	/* price multiplier for Cb/Cr (subdivide.c:166) */
   int	 rpf_mantissa;		/* lib/rpf.c: bits of the weight quantiser */
   float rpf_range;
   int	 dc_rpf_mantissa;
   float dc_rpf_range;
   int	 second_domain_block;	/* approx.c:103 (CLI -z 2) */
   int	 state_capacity;	/* workspace states per tile; 0 = pick from geometry */
} fb200_params_t;

/*
 *  Finished automaton of one tile, caller-allocated arrays of 'capacity' states
 *  (same meaning and element types as wfa_t, codec/wfa.h:112-138).  fb200_wfa_alloc()
 *  / fb200_wfa_free() are conveniences for C callers; foreign callers may fill the
 *  pointers themselves.
 */
typedef struct fb200_wfa
{
   int	     capacity;		/* in: states the arrays can hold */
   int	     status;		/* out: FB200_OK or the tile's own error */
   unsigned  states;		/* out */
   unsigned  basis_states;	/* out (3: the built-in "small.fco", input/basis.c:126) */
   unsigned  root_state;	/* out */
   float     costs [3];		/* out: subdivide() return value per band */
   float     err [3];		/* out: root range err / bits per band */
   float     tree_bits [3];
   float     matrix_bits [3];
   float     weights_bits [3];
   float    *final_distribution;	     /* [capacity] */
   uint8_t  *level_of_state;		     /* [capacity] */
   uint8_t  *domain_type;		     /* [capacity] */
   int16_t  *tree;			     /* [capacity][2] */
   uint16_t *x, *y;			     /* [capacity][2] */
   int16_t  *into;			     /* [capacity][2][6], -1 terminated */
   float    *weight;			     /* [capacity][2][6] */
   int16_t  *y_state;			     /* [capacity][2] */
   uint8_t  *y_column;		     /* [capacity][2] */
   /* predicted frames (fb200_encode_predicted; may be NULL): motion vectors of the ranges,
      mv_t of wfa->mv_tree, codec/wfa.h:62-71 -- type 0 none, 1 forward, 2 backward,
      3 interpolated */
   int8_t   *mv_type, *mv_fx, *mv_fy;	     /* [capacity][2] */
   int8_t   *mv_bx, *mv_by;		     /* [capacity][2] backward vectors (B frames) */
   /* in: smallest range level this tile starts with when > 0 (0: params.lc_min_level); out: the
      level after the frame.  The chroma bands of a colour frame raise c->options.lc_min_level for
      good (codec/coder.c:785-797), so the frames of a colour sequence hand it on. */
   int	     lc_min_level;
   /* out: the values the reference's percent meter prints while it codes band b
      (codec/subdivide.c:323-337), bit p of the 128-bit set = "p%" was shown */
   uint32_t  progress [3][4];
   /* colour frames through fb200_encode_predicted (may be NULL): in / out, [FB200_MAXSTATES][2].  The
      reference writes wfa->y_column of a colour frame's three virtual states to the stream
      (output/matrices.c:491-516) although nothing ever sets those entries: they hold what the last
      state with that number left there (append_transitions, codec/control.c:190-196; never cleared,
      codec/wfalib.c:276-310) -- in this frame or in any frame of the sequence before it.  The array
      is that history by the reference's state numbers: zeros before the first frame, then handed
      from every coded frame to the next in coding order. */
   uint8_t  *y_column_history;
} fb200_wfa_t;

/* one record per approximate_range() call (debug / parity tracing, optional) */
typedef struct fb200_trace_rec
{
   uint16_t level, image, address, x, y;
   int16_t  y_state;
   uint16_t states;
   int16_t  n_edges;		/* -1: costs >= max_costs (no approximation kept) */
   float    max_costs, price, costs, err, matrix_bits, weights_bits;
   int16_t  into [6];
   float    weight [6];
} fb200_trace_rec_t;

typedef struct fb200_ctx fb200_ctx_t;	/* opaque: device workspace for up to N tiles */

/* work / timing counters of the last launch */
typedef struct fb200_stats
{
   float    kernel_ms;		/* CUDA-event time of the tile kernel(s) */
   float    h2d_ms, d2h_ms;
   uint64_t h2d_bytes, d2h_bytes;
   uint64_t ip_bytes;		/* algorithmic bytes of the range x state products */
   uint64_t mp_calls, mp_steps, pass2, blocks, states;
   uint64_t mp_bytes;		/* algorithmic bytes of the pursuits: 8 D per call + 4 D per step */
   uint64_t ss_bytes;		/* algorithmic bytes of the state x state rows */
   uint64_t cyc_total, cyc_T, cyc_mp, cyc_append; /* SM cycles per phase, summed over tiles */
   uint64_t lap [16];		/* thread-0 cycles per sub-phase (profiling aid) */
   int	    kernel_launches;
} fb200_stats_t;

/*
 *  Work done by every launch of this process since the last reset, whichever entry point made it
 *  (fiasco_coder() of libfiasco included): for measurement harnesses that call the public API and
 *  still want the kernel time and the bytes that crossed the bus.
 */
typedef struct fb200_counters
{
   double   kernel_ms;		/* device time of the tile kernel launches (CUDA events) */
   uint64_t launches;		/* tile kernel launches, capacity retries included */
   uint64_t h2d_bytes, d2h_bytes;
} fb200_counters_t;
void fb200_counters_get (fb200_counters_t *out);
void fb200_counters_reset (void);


/* Fill 'p' from user-level options the way alloc_coder() does (coder.c:249-327).
   optimize follows the CLI: 0 => levels [6,10], 3 edges; 1 => [4,12], 5 edges;
   2 => 1 + second_domain_block; >= 3 => FB200_EUNSUPPORTED (full_search has
   undefined behaviour in the reference). */
int fb200_params_init (fb200_params_t *p, int width, int height, int bands,
		       float quality, int optimize, char *err, size_t errlen);

/* number of usable CUDA devices (0 if none / no driver) */
int fb200_device_count (void);

/* Create / destroy a device workspace for up to max_tiles tiles of geometry 'p' on
   CUDA device 'device'. */
int  fb200_create (fb200_ctx_t **ctx, const fb200_params_t *p, int max_tiles,
		   int device, char *err, size_t errlen);
void fb200_destroy (fb200_ctx_t *ctx);

/* Host-buffer path (what libfiasco's fiasco_coder() uses): planes[t * bands + b] is the
   b-th band of tile t in the reference's internal pixel format (int16 12.4 fixed
   point, lib/image.c:362,383-385), width*height each.  Copies in, runs, copies the
   automata out.  trace / trace_cap may be NULL / 0. */
int fb200_encode_tiles (fb200_ctx_t *ctx, int n_tiles,
			const int16_t *const *planes, fb200_wfa_t *out,
			fb200_trace_rec_t *trace, int trace_cap, int *trace_len,
			char *err, size_t errlen);

/* Device-resident path (benchmarks, callers that already hold frames in HBM):
   upload once, launch any number of times on a caller stream, download when needed. */
int fb200_upload (fb200_ctx_t *ctx, int n_tiles, const int16_t *const *planes,
		  char *err, size_t errlen);
int fb200_launch (fb200_ctx_t *ctx, int n_tiles, void *cuda_stream,
		  char *err, size_t errlen);
int fb200_download (fb200_ctx_t *ctx, int n_tiles, fb200_wfa_t *out,
		    fb200_trace_rec_t *trace, int trace_cap, int *trace_len,
		    char *err, size_t errlen);
int fb200_sync (fb200_ctx_t *ctx, char *err, size_t errlen);
void fb200_get_stats (const fb200_ctx_t *ctx, fb200_stats_t *stats);
/* tiles the device keeps resident at once (SM count x thread blocks per SM): the batch
   size that fills one wave */
int fb200_resident_tiles (const fb200_ctx_t *ctx);
int fb200_state_capacity (const fb200_ctx_t *ctx);

/* helpers for C callers */
int  fb200_wfa_alloc (fb200_wfa_t *wfa, int capacity);
void fb200_wfa_free (fb200_wfa_t *wfa);

/* pure-function probes (known-answer tests of the device arithmetic): evaluate n values
   on the GPU.  kind: 0 rtob(f[i]; mantissa a[i], range_e b[i]) -> out_i
		      1 btor(a[i]; mantissa b[i], range_e c[i]) -> out_f
		      2 bits_bin_code(a[i], b[i]) -> out_i
		      3 -log2(a[i] / (float) b[i]) as fp32 -> out_f  (tree/aac/rle rate terms) */
int fb200_probe (int kind, int n, const float *f, const int *a, const int *b,
		 const int *c, int *out_i, float *out_f, char *err, size_t errlen);

/*
 *  Norms tables of the motion search for a predicted frame (replaces the lazy, per-block
 *  fill_norms_table() of the reference, codec/mwfa.c:544-602, by one launch over all blocks): for
 *  every block of bintree level 'level' that lies completely inside the frame and every
 *  displacement (mx, my) in [-search_range, search_range)^2 the squared norm of
 *  (original - reference displaced by (mx, my)) / 16, bit-identical to the reference's sums.
 *  orig / past: width * height shorts (the coder's pixel format, host memory).  norms (host):
 *  [rows of blocks][columns of blocks][(my + sr) * 2 sr + (mx + sr)], 0 for blocks or
 *  displacements that leave the frame.  kernel_ms (or NULL): device time of the kernel.  A
 *  stand-alone kernel (analysis, benchmarks): fb200_encode_predicted() computes its tables inside
 *  the tile kernel in the reference's lazy order, because the partial sums that aborted subtrees
 *  leave behind are part of the reference's result.
 */
int fb200_motion_norms (int device, const int16_t *orig, const int16_t *past, int width,
			int height, int level, int search_range, float *norms, float *kernel_ms,
			char *err, size_t errlen);

/*
 *  Predicted frames of a sequence (replaces, in the reference coder, the same subdivide() call
 *  at codec/coder.c:743 when the frame is a P or B frame, i.e. with its third alternative:
 *  predict_range / mc_prediction, codec/prediction.c:96,262; find_P_frame_mc, find_B_frame_mc
 *  (without the cross-B search, which the reference ties to the half-pixel flag, coder.c:359),
 *  fill_norms_table, find_best_mv, codec/mwfa.c:301,341,544,686; and the nested subdivide() over the prediction error
 *  with the delta pool and delta coefficient model).  Levels and search range as in
 *  c_options_t (codec/options.h: p_min_level, p_max_level, search_range; CLI defaults 6, 10, 16).
 */
#define FB200_FRAME_INTRA 4	/* an intra frame without any prediction on the workspace of predicted frames:
				   the intra frames of a COLOUR sequence that has predicted frames, so that all
				   its frames keep fb200_wfa_t.y_column_history */
#define FB200_FRAME_ND 3	/* an INTRA frame coded with nondeterministic prediction (`cfiasco
				   --prediction', fiasco_c_options_set_prediction; nd_prediction,
				   codec/prediction.c:371): the third alternative of subdivide() is the
				   range's DC component plus a nested pass over the difference.  No
				   reference frames: 'past' and 'future' may be NULL */

typedef struct fb200_motion
{
   int frame_type;		/* 1 = P frame, 2 = B frame, FB200_FRAME_ND */
   int p_min_level, p_max_level;
   int search_range;		/* vectors in [-search_range, search_range), full pixel */
} fb200_motion_t;

/* A device workspace for predicted frames of geometry 'p' (one independent sequence per tile). */
int fb200_create_predicted (fb200_ctx_t **ctx, const fb200_params_t *p,
			    const fb200_motion_t *motion, int max_tiles, int device,
			    char *err, size_t errlen);

/*
 *  Encode one predicted frame per tile: planes[t] is the frame, past[t] the REGENERATED previous
 *  frame of the same sequence (future[t]: the regenerated next reference frame, B frames only, else
 *  NULL) (fiasco_regenerate_frame() of include/fiasco_host.h; what
 *  decode_image + restore_mc leave in the reference, codec/coder.c:642-651), both width*height
 *  shorts in host memory.  The automata come back as the device leaves them: states of losing
 *  alternatives are holes (level_of_state == 255) and out[t].mv_* hold the vectors;
 *  fiasco_finish_predicted_frame() closes the holes and derives the delta flags.
 *
 *  Colour sequences (params.bands == 3): planes holds three pointers per tile (Y, Cb, Cr), past[t] /
 *  future[t] point at the three regenerated planes of the reference frame behind each other
 *  (fiasco_regenerate_colour_frame()).  The luminance band is coded with prediction, the luminance
 *  tree's motion compensation is then taken off the chroma planes on the device (subtract_mc,
 *  codec/mwfa.c:156-299) and Cb, Cr are coded without prediction (codec/coder.c:757-849).  The frames
 *  of a colour sequence are chained: hand out[t].lc_min_level and out[t].y_column_history of the
 *  frame coded before to the next call (fb200_wfa_t).  Contexts of type FB200_FRAME_ND and
 *  FB200_FRAME_INTRA code intra frames: past and future may be NULL.
 */
int fb200_encode_predicted (fb200_ctx_t *ctx, int n_tiles, const int16_t *const *planes,
			    const int16_t *const *past, const int16_t *const *future,
			    fb200_wfa_t *out, char *err, size_t errlen);

const char *fb200_version (void);

#ifdef __cplusplus
}
#endif
#endif /* FIASCO_B200_H */
