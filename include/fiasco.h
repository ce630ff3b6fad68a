/*
 *  fiasco.h -- public C API of libfiasco as provided by the B200-native implementation.
 *
 *  Re-declared from the reference's public header (/root/reference/fiasco.h:39-398) and its
 *  man pages (doc/fiasco_coder.3, doc/fiasco_options.3): same names, same argument meaning,
 *  same struct layouts, same error convention (functions return 1 on success and 0 on
 *  failure, the text is available from fiasco_get_error_message()), so that the reference
 *  command line front end bin/cwfa.c links against this library unchanged.
 *
 *  Scope: the CODER side of the API.  The decoder / image / renderer classes of the
 *  reference header (fiasco_decoder_*, fiasco_image_*, fiasco_renderer_*,
 *  fiasco_d_options_*) are not part of the accelerated path and are not provided; use the
 *  reference's dfiasco to decode the streams written here (they are byte-identical to the
 *  reference coder's).
 */
#ifndef _FIASCO_H
#define _FIASCO_H 1

#ifdef __cplusplus
extern "C" {
#endif

/* verbosity of messages on stderr (reference fiasco.h:52-54) */
typedef enum {FIASCO_NO_VERBOSITY,
	      FIASCO_SOME_VERBOSITY,
	      FIASCO_ULTIMATE_VERBOSITY} fiasco_verbosity_e;

/* image tiling methods (reference fiasco.h:67-70); accepted and -- exactly like the
   reference encoder at this commit -- without effect on the stream */
typedef enum {FIASCO_TILING_SPIRAL_ASC,
	      FIASCO_TILING_SPIRAL_DSC,
	      FIASCO_TILING_VARIANCE_ASC,
	      FIASCO_TILING_VARIANCE_DSC} fiasco_tiling_e;

/* range of the reduced precision format (reference fiasco.h:79-82) */
typedef enum {FIASCO_RPF_RANGE_0_75,
	      FIASCO_RPF_RANGE_1_00,
	      FIASCO_RPF_RANGE_1_50,
	      FIASCO_RPF_RANGE_2_00} fiasco_rpf_range_e;

/* progress meter (reference fiasco.h:90-92) */
typedef enum {FIASCO_PROGRESS_NONE,
	      FIASCO_PROGRESS_BAR,
	      FIASCO_PROGRESS_PERCENT} fiasco_progress_e;

/* class of advanced coder options (reference fiasco.h:132-174): 12 methods + private */
typedef struct fiasco_c_options
{
   void (*delete)	     (struct fiasco_c_options *options);
   int (*set_tiling)	     (struct fiasco_c_options *options,
			      fiasco_tiling_e method, unsigned exponent);
   int (*set_frame_pattern)  (struct fiasco_c_options *options, const char *pattern);
   int (*set_basisfile)	     (struct fiasco_c_options *options, const char *filename);
   int (*set_chroma_quality) (struct fiasco_c_options *options, float quality_factor,
			      unsigned dictionary_size);
   int (*set_optimizations)  (struct fiasco_c_options *options,
			      unsigned min_block_level, unsigned max_block_level,
			      unsigned max_elements, unsigned dictionary_size,
			      unsigned optimization_level);
   int (*set_prediction)     (struct fiasco_c_options *options, int intra_prediction,
			      unsigned min_block_level, unsigned max_block_level);
   int (*set_video_param)    (struct fiasco_c_options *options,
			      unsigned frames_per_second, int half_pixel_prediction,
			      int cross_B_search, int B_as_past_ref);
   int (*set_quantization)   (struct fiasco_c_options *options, unsigned mantissa,
			      fiasco_rpf_range_e range, unsigned dc_mantissa,
			      fiasco_rpf_range_e dc_range);
   int (*set_progress_meter) (struct fiasco_c_options *options, fiasco_progress_e type);
   int (*set_smoothing)	     (struct fiasco_c_options *options, int smoothing);
   int (*set_comment)	     (struct fiasco_c_options *options, const char *comment);
   int (*set_title)	     (struct fiasco_c_options *options, const char *title);
   void *private;
} fiasco_c_options_t;

/* miscellaneous (reference fiasco.h:220-222) */
const char *fiasco_get_error_message (void);
void fiasco_set_verbosity (fiasco_verbosity_e level);
fiasco_verbosity_e fiasco_get_verbosity (void);

/*
 *  Encode the image / video frames named by the NULL terminated array 'inputname' (raw
 *  PGM / PPM; "-" or NULL = stdin; templates "prefix[start-end{+,-}step]suffix") into the
 *  FIASCO stream 'outputname' ("-" or NULL = stdout).  quality: 1 (worst) .. 100 (best).
 *  options == NULL selects the library defaults.  Returns 1 on success, 0 otherwise.
 *  (reference fiasco.h:303-306, codec/coder.c:85)
 */
int fiasco_coder (char const * const *inputname, const char *outputname,
		  float quality, const fiasco_c_options_t *options);

/* coder options (reference fiasco.h:312-398, codec/options.c) */
fiasco_c_options_t *fiasco_c_options_new (void);
void fiasco_c_options_delete (fiasco_c_options_t *options);
int fiasco_c_options_set_smoothing (fiasco_c_options_t *options, int smoothing);
int fiasco_c_options_set_frame_pattern (fiasco_c_options_t *options, const char *pattern);
int fiasco_c_options_set_tiling (fiasco_c_options_t *options, fiasco_tiling_e method,
				 unsigned exponent);
int fiasco_c_options_set_basisfile (fiasco_c_options_t *options, const char *filename);
int fiasco_c_options_set_chroma_quality (fiasco_c_options_t *options, float quality_factor,
					 unsigned dictionary_size);
int fiasco_c_options_set_optimizations (fiasco_c_options_t *options,
					unsigned min_block_level, unsigned max_block_level,
					unsigned max_elements, unsigned dictionary_size,
					unsigned optimization_level);
int fiasco_c_options_set_prediction (fiasco_c_options_t *options, int intra_prediction,
				     unsigned min_block_level, unsigned max_block_level);
int fiasco_c_options_set_video_param (fiasco_c_options_t *options,
				      unsigned frames_per_second, int half_pixel_prediction,
				      int cross_B_search, int B_as_past_ref);
int fiasco_c_options_set_quantization (fiasco_c_options_t *options, unsigned mantissa,
				       fiasco_rpf_range_e range, unsigned dc_mantissa,
				       fiasco_rpf_range_e dc_range);
int fiasco_c_options_set_progress_meter (fiasco_c_options_t *options, fiasco_progress_e type);
int fiasco_c_options_set_comment (fiasco_c_options_t *options, const char *comment);
int fiasco_c_options_set_title (fiasco_c_options_t *options, const char *title);

#ifdef __cplusplus
}
#endif
#endif /* not _FIASCO_H */
