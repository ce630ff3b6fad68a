/*
 *  fiasco_oracle.h -- CPU restatement of the FIASCO encoder hot path.
 *
 *  TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 *  cpu_baseline / --impl reference legs may build, load or call this.  The product
 *  (fiasco_b200/) never includes, links or falls back to anything in oracle/.
 *
 *  Parity status: PINNED -- against the reference itself, which compiles here
 *  (oracle/_ref, built in place from /root/reference by oracle/Makefile): the WFA this
 *  restatement produces is compared state-for-state / edge-for-edge / bit-for-bit with
 *  the reference's own output (oracle/_ref/wfadump of the reference .fco) and its
 *  per-range trace with the reference's (oracle/_ref/seamdump, ld --wrap on the
 *  unmodified objects); the golden copies live in tests/golden/.
 */
#ifndef FIASCO_ORACLE_H
#define FIASCO_ORACLE_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FO_MAXEDGES  5		/* codec/wfa.h:20 */
#define FO_MAXSTATES 6000	/* codec/wfa.h:21 */
#define FO_MAXLABELS 2		/* codec/wfa.h:22 */
#define FO_MAXLEVEL  22		/* codec/wfa.h:23 */

/* coder options after the CLI / fiasco_c_options_* layer, before the clamping of
   codec/coder.c:260-296 (fo_encode applies that clamping itself) */
typedef struct fo_params
{
   int	 width, height;		/* image geometry (even numbers) */
   int	 color;			/* 0: one grey band, 1: Y, Cb, Cr (4:4:4) */
   float quality;		/* price = 128 * 64 / quality (coder.c:164) */
   int	 lc_min_level;		/* options.c:67-74 / cwfa.c:332-345 */
   int	 lc_max_level;
   int	 images_level;		/* 5 */
   int	 max_elements;		/* edges per linear combination */
   int	 max_states;		/* dictionary size */
   int	 chroma_max_states;	/* 40 */
   float chroma_decrease;	/* 2.0 */
   int	 rpf_mantissa;		/* 3 */
   int	 rpf_range_e;		/* fiasco_rpf_range_e: 0=.75 1=1.0 2=1.5 3=2.0 */
   int	 dc_rpf_mantissa;	/* 5 */
   int	 dc_rpf_range_e;	/* 1 */
   int	 second_domain_block;	/* optimisation level >= 1 */
   int	 check_for_underflow;	/* optimisation level >= 2 */
   int	 check_for_overflow;	/* optimisation level >= 2 */
   int	 full_search;		/* optimisation level >= 2 (UB in the reference) */
} fo_params_t;

/* the finished automaton (codec/wfa.h:112-138, the fields the still-image path fills) */
typedef struct fo_wfa
{
   unsigned states, basis_states, root_state;
   unsigned level;		/* image level (wfainfo->level) */
   float    final_distribution [FO_MAXSTATES];
   uint8_t  level_of_state [FO_MAXSTATES];
   uint8_t  domain_type [FO_MAXSTATES];
   int16_t  tree [FO_MAXSTATES][FO_MAXLABELS];
   uint16_t x [FO_MAXSTATES][FO_MAXLABELS];
   uint16_t y [FO_MAXSTATES][FO_MAXLABELS];
   int16_t  into [FO_MAXSTATES][FO_MAXLABELS][FO_MAXEDGES + 1];
   float    weight [FO_MAXSTATES][FO_MAXLABELS][FO_MAXEDGES + 1];
   int16_t  y_state [FO_MAXSTATES][FO_MAXLABELS];
   uint8_t  y_column [FO_MAXSTATES][FO_MAXLABELS];
   /* root range summary per band (what print_statistics reports, coder.c:894) */
   float    costs [3], err [3], tree_bits [3], matrix_bits [3], weights_bits [3];
   /* predicted frames (codec/wfa.h:62-71,126,137): motion compensation of the range
      [state][label] -- type 0 none, 1 forward, 2 backward, 3 interpolated -- and the
      states that describe a prediction error instead of image content */
   int8_t   mv_type [FO_MAXSTATES][FO_MAXLABELS];
   int8_t   mv_fx [FO_MAXSTATES][FO_MAXLABELS], mv_fy [FO_MAXSTATES][FO_MAXLABELS];
   int8_t   mv_bx [FO_MAXSTATES][FO_MAXLABELS], mv_by [FO_MAXSTATES][FO_MAXLABELS];
   uint8_t  delta_state [FO_MAXSTATES];
   int	    frame_type;		/* 0 intra, 1 predicted, 2 bidirectional */
   int	    frame_number;	/* display number (sequences are coded out of display order) */
} fo_wfa_t;

/* work counters */
typedef struct fo_stats
{
   uint64_t subdivide_calls, mp_calls, pass1, pass2, ortho_steps, accepted,
	    append_states, leaf_dots, ipss_lookups, mp_domains;
   uint64_t ip_bytes;		/* algorithmic bytes of the range x state kernel:
				   sum over lc_max blocks of 4*2^lc_max + 376*S_b */
   uint64_t blocks;
} fo_stats_t;

void fo_default_params (fo_params_t *p, int width, int height, int color,
			float quality, int optimize /* CLI -z, 0..3 */);

/*
 *  Encode one still frame.  planes[b] are the frame's bands in the reference's internal
 *  pixel format (lib/image.c:362,383-385: 12.4 fixed point shorts, row-major,
 *  width*height each; 1 band grey, 3 bands Y Cb Cr).  Returns 0 on success, else an
 *  error code (message in errbuf).  If trace != NULL a per-seam trace in the format of
 *  oracle/seamdump.c is written.
 */
int fo_encode (const fo_params_t *p, const int16_t *const planes [3],
	       fo_wfa_t *out, fo_stats_t *stats, FILE *trace,
	       char *errbuf, size_t errlen);

/* pixel format conversion of lib/image.c:352-386 (8-bit PNM samples -> shorts) */
void fo_grey_to_plane (const uint8_t *grey, size_t n, int16_t *plane);
void fo_rgb_to_planes (const uint8_t *rgb, size_t n, int16_t *y, int16_t *cb,
		       int16_t *cr);

/* small pure functions of the path, exported for known-answer tests */
int	 fo_rtob (float f, unsigned mantissa_bits, int range_e);  /* lib/rpf.c:59 */
float	 fo_btor (int b, unsigned mantissa_bits, int range_e);	  /* lib/rpf.c:113 */
unsigned fo_bits_bin_code (unsigned value, unsigned maxval);	  /* lib/misc.c:296 */
float	 fo_neg_log2f (int count, int total);	/* -log2 (count / (real_t) total) */
void	 fo_tree_model_kat (unsigned level, unsigned *counts, unsigned *total,
			    float *child_bits, float *leaf_bits); /* bintree.c:55,70 */
unsigned fo_image_level (unsigned width, unsigned height);	  /* coder.c:249-256 */

/*
 *  Regenerate the frame an automaton describes (decode_image, codec/decoder.c:412, 4:4:4): what
 *  the coder itself does after every frame of a video (codec/coder.c:647).  planes [b]:
 *  width * height shorts per band in the reference's pixel format.  Returns 0 on success.
 */
int fo_decode_image (const fo_wfa_t *wfa, int color, unsigned width, unsigned height,
		     int16_t *const planes [3]);

/*
 *  Build an automaton from the canonical text of ONE frame ("s" / "e" / "m" / "d" lines of
 *  oracle/wfadump.c): the basis states as in input/basis.c:126-131, transitions and motion
 *  vectors from the text, final distributions recomputed (codec/wfalib.c:154).  Used to pin the
 *  decoder side on automata the reference produced.  Returns 0 on success.
 */
int fo_wfa_from_dump (const char *text, unsigned root_state, fo_wfa_t *wfa);

/*
 *  restore_mc (codec/motion.c:37-230) for a grey predicted frame: add the motion compensated
 *  blocks of the regenerated previous frame 'past' to 'image' (both width * height shorts).
 */
void fo_restore_mc (const fo_wfa_t *wfa, unsigned width, unsigned height, int half_pixel,
		    int16_t *image, const int16_t *past, const int16_t *future);
/* the same for a colour frame (three planes each): the luminance tree's vectors move all bands, the
   chroma bands are clipped to 8 bits afterwards (motion.c:59-62, 192-224) */
void fo_restore_mc_colour (const fo_wfa_t *wfa, unsigned width, unsigned height, int16_t *image,
			   const int16_t *past, const int16_t *future);

/*
 *  Encode a sequence (video_coder / frame_coder, codec/coder.c:490-892): frame 0 intra, the
 *  others I, P or B by 'pattern'; predicted frames use motion compensation against the regenerated
 *  reference frames (codec/prediction.c, codec/mwfa.c).  Grey: frames [f], width * height shorts.
 *  Colour (p->color): frames [3 f + b], b = Y, Cb, Cr; the luminance band is coded with prediction,
 *  the luminance tree's motion compensation is taken off the chroma planes (subtract_mc,
 *  codec/mwfa.c:156) before Cb and Cr are coded without prediction (coder.c:757-849).
 *  out [f]: automaton of the f-th CODED frame; reconst (or NULL): the regenerated frames (colour:
 *  three planes each).  0 on success.
 */
int fo_encode_video (const fo_params_t *p, int n_frames, const int16_t *const *frames,
		     const char *pattern, int p_min_level, int p_max_level, int search_range,
		     fo_wfa_t *out, int16_t *reconst, FILE *trace, char *errbuf, size_t errlen);

/* fill_norms_table (codec/mwfa.c:544-602) for the block at (x0, y0) of bintree level 'level':
   out [(my + sr) * 2 sr + (mx + sr)] */
void fo_fill_norms_table (const int16_t *orig, const int16_t *past, unsigned width, unsigned height,
			  unsigned x0, unsigned y0, unsigned level, unsigned search_range, float *out);

/* `cfiasco --prediction' (fiasco_c_options_set_prediction): the intra frames of fo_encode_video try
   the nondeterministic prediction of codec/prediction.c:371 (DC component + delta image) */
void fo_set_nd_prediction (int on);
/* test aid: DC weights that rounded to zero (infinitely many bits in the reference, codec/coeff.c:237 reading in
   front of its table: never taken) since nondeterministic prediction was last switched on */
unsigned fo_nd_zero_weights (void);

/* design check of the device's state handling for predicted frames (see fiasco_oracle.c) */
void fo_set_holes_mode (int on);
void fo_close_holes (fo_wfa_t *wfa);

/* canonical text dump, same grammar as oracle/wfadump.c ("s"/"e" lines of one frame) */
void fo_dump_wfa (const fo_wfa_t *wfa, const fo_params_t *p, FILE *f);

#ifdef __cplusplus
}
#endif
#endif
