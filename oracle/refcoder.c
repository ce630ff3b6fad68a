/*
 *  refcoder.c -- TEST INFRASTRUCTURE: calls the UNMODIFIED reference library's fiasco_coder()
 *  (linked from oracle/_ref/libfiasco_ref.a) with options the reference command line front end
 *  parses but never passes on (bin/cwfa.c has no call of fiasco_c_options_set_video_param: its
 *  --fps, --half-pixel, --cross-B-search and --B-as-past-ref flags have no effect).  Our own
 *  code; contains no reference source text.
 *
 *      refcoder out.fco quality pattern fps half_pixel cross_B B_as_past_ref frame...
 *
 *  Everything else as the CLI sets it for its defaults (bin/cwfa.c:253-388).
 */
#include <stdio.h>
#include <stdlib.h>
#include "fiasco.h"

int
main (int argc, char **argv)
{
   fiasco_c_options_t *o;

   if (argc < 9)
   {
      fprintf (stderr, "usage: refcoder out.fco quality pattern fps half_pixel cross_B B_as_past_ref frame...\n");
      return 2;
   }
   o = fiasco_c_options_new ();
   fiasco_c_options_set_frame_pattern (o, argv [3]);
   fiasco_c_options_set_chroma_quality (o, 2.0, 40);
   fiasco_c_options_set_smoothing (o, 70);
   fiasco_c_options_set_progress_meter (o, FIASCO_PROGRESS_NONE);
   fiasco_c_options_set_tiling (o, FIASCO_TILING_VARIANCE_DSC, 4);
   fiasco_c_options_set_optimizations (o, 6, 10, 3, 10000, 0);
   fiasco_c_options_set_prediction (o, 0, 6, 10);
   fiasco_c_options_set_quantization (o, 3, FIASCO_RPF_RANGE_1_50, 5, FIASCO_RPF_RANGE_1_00);
   if (!fiasco_c_options_set_video_param (o, (unsigned) atoi (argv [4]), atoi (argv [5]), atoi (argv [6]),
					  atoi (argv [7])))
   {
      fprintf (stderr, "refcoder: %s\n", fiasco_get_error_message ());
      return 1;
   }
   fiasco_set_verbosity (FIASCO_NO_VERBOSITY);
   if (!fiasco_coder ((char const *const *) (argv + 8), argv [1], (float) atof (argv [2]), o))
   {
      fprintf (stderr, "refcoder: %s\n", fiasco_get_error_message ());
      return 1;
   }
   fiasco_c_options_delete (o);
   return 0;
}
