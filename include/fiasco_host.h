/*
 *  fiasco_host.h -- host-side helpers of libfiasco (B200 build) that sit between the C ABI
 *  of the GPU path (fiasco_b200.h) and the FIASCO stream format: serialise automata that
 *  fb200_encode_tiles() returned.  fiasco_coder() uses exactly this internally; callers
 *  that drive the GPU path themselves (tile-split mode: one independent FIASCO stream per
 *  tile) use it to obtain ordinary .fco files that the reference dfiasco
 *  decodes.
 */
#ifndef FIASCO_HOST_H
#define FIASCO_HOST_H

#include "fiasco_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* stream-level parameters that are not part of the automaton (reference wfa_info_t,
   codec/wfa.h:86-110, as far as an intra stream needs them) */
typedef struct fiasco_stream_info
{
   int	       width, height, color;
   unsigned    max_states;		/* as written in the header (min (dictionary, 6000)) */
   unsigned    chroma_max_states;
   unsigned    p_min_level, p_max_level;
   unsigned    smoothing;		/* 70 */
   unsigned    fps;			/* 25 */
   int	       rpf_mantissa, rpf_range_e, dc_rpf_mantissa, dc_rpf_range_e;
   const char *title, *comment;		/* may be NULL */
   int	       nd_prediction;		/* coded with `--prediction' (fiasco_c_options_set_prediction): every
					   frame carries its nondeterminism tree (output/write.c:100, output/nd.c) */
} fiasco_stream_info_t;

/* defaults of the reference command line front end for a stream coded with params 'p'
   (bin/cwfa.c:36-90, codec/coder.c:285-296) */
void fiasco_stream_info_init (fiasco_stream_info_t *info, const fb200_params_t *p);

/*
 *  Write n_frames intra automata (one per frame, same geometry) as one FIASCO stream to
 *  'filename' ("-" = stdout; searched like the coder's output name).  Returns 1 on
 *  success, 0 on failure (text from fiasco_get_error_message()).  The bytes equal what
 *  the reference coder writes for the same automata (output/write.c:53).
 */
int fiasco_write_stream (const char *filename, const fiasco_stream_info_t *info,
			 const fb200_wfa_t *frames, int n_frames);

/*
 *  Predicted frames (reference: mv_t / delta_state of wfa_t, codec/wfa.h:62-71,126,137): what a
 *  predicted frame's automaton carries beside the fields of fb200_wfa_t.  Arrays are [states][2]
 *  (mv_*) and [states] (delta_state); for intra frames the pointers may be NULL.
 */
typedef struct fiasco_frame_motion
{
   int		  frame_type;	/* 0 intra, 1 predicted (P), 2 bidirectional (B) */
   int		  frame_number;	/* display number: sequences with B frames are coded out of order */
   const int8_t	 *mv_type;	/* 0 none, 1 forward, 2 backward, 3 interpolated */
   const int8_t	 *mv_fx, *mv_fy;
   const int8_t	 *mv_bx, *mv_by;	/* backward vectors (B frames), else may be NULL */
   const uint8_t *delta_state;	/* states that describe a prediction error (also of an intra frame coded
				   with nondeterministic prediction) */
} fiasco_frame_motion_t;

/*
 *  fiasco_write_stream() for sequences with predicted frames: writes the motion tree and the
 *  vectors (output/mc.c:75) and codes the weights of delta states in their own contexts
 *  (output/weights.c:38).  motion == NULL: all frames intra.  The automata of predicted frames
 *  come from fb200_encode_predicted() + fiasco_finish_predicted_frame(); frames are passed in
 *  coding order (fiasco_frame_motion_t.frame_number is the display number).
 */
int fiasco_write_video_stream (const char *filename, const fiasco_stream_info_t *info,
			       const fb200_wfa_t *frames, const fiasco_frame_motion_t *motion,
			       int n_frames, unsigned search_range);

/*
 *  Regenerate the grey frame an automaton describes, in the coder's pixel format (shorts, 12.4
 *  fixed point), the way the reference coder does after every frame of a sequence (decode_image,
 *  codec/decoder.c:412; restore_mc, codec/motion.c:37, when 'motion' says the frame is predicted
 *  from 'past' and, for B frames, 'future').  out / past / future: width * height shorts.  Returns 1 on
 *  success, 0 on failure.
 */
int fiasco_regenerate_frame (const fb200_wfa_t *wfa, const fiasco_frame_motion_t *motion,
			     int width, int height, const int16_t *past, const int16_t *future,
			     int16_t *out);

/*
 *  The same for a colour frame (4:4:4, as the coder regenerates it): out / past / future hold three
 *  planes Y, Cb, Cr of width * height shorts behind each other.  The vectors of the luminance tree
 *  move all three bands and the chroma bands of a predicted frame are clipped to 8 bits afterwards
 *  (restore_mc, codec/motion.c:59-62, 192-224).
 */
int fiasco_regenerate_colour_frame (const fb200_wfa_t *wfa, const fiasco_frame_motion_t *motion,
				    int width, int height, const int16_t *past, const int16_t *future,
				    int16_t *out);

/*
 *  Finish the automaton of a predicted frame the way the device leaves it:
 *  close the holes of losing split alternatives (states marked level_of_state == 255) by a monotone
 *  renumbering and derive the delta flags from the structure (locate_delta_images,
 *  codec/wfalib.c:699, called at codec/coder.c:876).  In place; mv_* are [states][2], delta_state
 *  [states].  Returns the new number of states, 0 on failure.
 */
int fiasco_finish_predicted_frame (fb200_wfa_t *wfa, int8_t *mv_type, int8_t *mv_fx, int8_t *mv_fy,
				   int8_t *mv_bx, int8_t *mv_by, uint8_t *delta_state);

#ifdef __cplusplus
}
#endif
#endif
