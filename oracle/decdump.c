/*
 *  decdump.c -- TEST INFRASTRUCTURE (oracle side), not part of the product.
 *
 *  Regenerates every frame of a FIASCO stream with the UNMODIFIED reference functions the
 *  coder itself uses for its reference frames (codec/coder.c:647-651): decode_image()
 *  (codec/decoder.c:412) and, for predicted frames, restore_mc() (codec/motion.c:37) against
 *  the previously regenerated frame -- and writes the frames' pixels in the reference's
 *  internal format (shorts, 12.4 fixed point, lib/image.h) so that a restatement of the
 *  decoder can be pinned bit for bit.  Linked against oracle/_ref/libfiasco_ref.a; contains no
 *  reference source text.
 *
 *  usage: decdump file.fco out.raw     (all frames, all bands, int16 little endian, appended)
 *         prints one line per frame: "frame <n> <type> <width> <height> <bands>"
 */
#include "config.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "types.h"
#include "macros.h"
#include "error.h"
#include "wfa.h"
#include "wfalib.h"
#include "bit-io.h"
#include "read.h"
#include "image.h"
#include "decoder.h"
#include "motion.h"
#include "fiasco.h"

int
main (int argc, char **argv)
{
   if (argc != 3)
   {
      fprintf (stderr, "usage: %s file.fco out.raw\n", argv [0]);
      return 2;
   }
   fiasco_set_verbosity (FIASCO_NO_VERBOSITY);
   try
   {
      wfa_t	*wfa   = alloc_wfa (NO);
      bitfile_t *input = open_wfa (argv [1], wfa->wfainfo);
      FILE	*out   = fopen (argv [2], "wb");
      image_t	*past = NULL, *future = NULL, *reconst = NULL;
      bool_t	 future_frame = NO;	/* the previous frame of the stream was a future reference */
      unsigned	 seen [4096], expected = 0;
      unsigned	 f;

      memset (seen, 0, sizeof seen);

      if (!out)
	 error ("cannot write %s", argv [2]);
      read_basis (wfa->wfainfo->basis_name, wfa);
      for (f = 0; f < wfa->wfainfo->frames; f++)
      {
	 unsigned  number = read_next_wfa (wfa, input);
	 image_t  *frame  = decode_image (wfa->wfainfo->width, wfa->wfainfo->height,
					  FORMAT_4_4_4, NULL, wfa);
	 unsigned  bands  = frame->color ? 3 : 1, b;

	 /* reference frames exactly as video_coder() keeps them (codec/coder.c:560-627): frames
	    arrive in coding order, a frame that is ahead of the display order is the future
	    reference of the B frames that follow */
	 if (wfa->frame_type == I_FRAME)
	 {
	    if (past) free_image (past);
	    if (future) free_image (future);
	    if (reconst) free_image (reconst);
	    past = future = reconst = NULL;
	 }
	 else if (wfa->frame_type == P_FRAME)
	 {
	    if (past) free_image (past);
	    past    = reconst;
	    reconst = NULL;
	    if (future) free_image (future);
	    future = NULL;
	 }
	 else if (future_frame)
	 {
	    if (future) free_image (future);
	    future  = reconst;
	    reconst = NULL;
	 }
	 else if (wfa->wfainfo->B_as_past_ref == YES)
	 {
	    if (past) free_image (past);
	    past    = reconst;
	    reconst = NULL;
	 }
	 else
	 {
	    if (reconst) free_image (reconst);
	    reconst = NULL;
	 }
	 if (number < 4096)
	    seen [number] = 1;
	 future_frame = number > expected;
	 while (expected < 4096 && seen [expected])
	    expected++;
	 if (wfa->frame_type != I_FRAME)
	    restore_mc (0, frame, past, future, wfa);
	 printf ("frame %u %d %u %u %u\n", number, (int) wfa->frame_type,
		 frame->width, frame->height, bands);
	 for (b = 0; b < bands; b++)
	    fwrite (frame->pixels [frame->color ? b : GRAY], sizeof (word_t),
		    (size_t) frame->width * frame->height, out);
	 reconst = frame;
	 remove_states (wfa->basis_states, wfa);
      }
      fclose (out);
      close_bitfile (input);
      return 0;
   }
   catch
   {
      fprintf (stderr, "decdump: %s\n", fiasco_get_error_message ());
      return 1;
   }
}
