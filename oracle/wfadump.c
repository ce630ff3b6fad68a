/*
 *  wfadump.c -- TEST INFRASTRUCTURE (oracle side), not part of the product.
 *
 *  Reads a FIASCO stream with the REFERENCE's own reader (open_wfa / read_basis /
 *  read_next_wfa, used the same way by the reference's bin/twfa.c:434-437) and prints
 *  the automaton of every frame in a canonical text form.  Two .fco files hold the same
 *  WFA iff their dumps are identical, so this pins "identical tree / domain indices /
 *  quantised weights" without needing byte-identical entropy coding.
 *
 *  This file contains no reference source text; it only calls reference functions and is
 *  linked against oracle/_ref/libfiasco_ref.a (built in place from /root/reference).
 *
 *  Output grammar (one record per line):
 *    info <width> <height> <level> <color> <frames> <release> <basis_states>
 *         <rpf_mantissa> <rpf_range_e> <dc_rpf_mantissa> <dc_rpf_range_e>
 *    frame <number> <frame_type> <states> <root_state>
 *    s <state> <level_of_state> <tree0> <tree1> <x0> <y0> <x1> <y1> <pred0> <pred1>
 *    e <state> <label> <into> <weight bits, hex> <weight %.9g>
 *    end
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
#include "fiasco.h"

static unsigned
float_bits (float f)
{
   unsigned u;
   memcpy (&u, &f, sizeof u);
   return u;
}

int
main (int argc, char **argv)
{
   if (argc != 2)
   {
      fprintf (stderr, "usage: %s file.fco\n", argv [0]);
      return 2;
   }
   fiasco_set_verbosity (FIASCO_NO_VERBOSITY);
   try
   {
      wfa_t     *wfa   = alloc_wfa (NO);
      bitfile_t *input = open_wfa (argv [1], wfa->wfainfo);
      unsigned   f;

      read_basis (wfa->wfainfo->basis_name, wfa);
      printf ("info %u %u %u %d %u %u %u %u %d %u %d\n",
	      wfa->wfainfo->width, wfa->wfainfo->height, wfa->wfainfo->level,
	      (int) wfa->wfainfo->color, wfa->wfainfo->frames,
	      wfa->wfainfo->release, wfa->basis_states,
	      wfa->wfainfo->rpf->mantissa_bits, (int) wfa->wfainfo->rpf->range_e,
	      wfa->wfainfo->dc_rpf->mantissa_bits,
	      (int) wfa->wfainfo->dc_rpf->range_e);
      for (f = 0; f < wfa->wfainfo->frames; f++)
      {
	 unsigned number = read_next_wfa (wfa, input);
	 unsigned state, label, edge;

	 printf ("frame %u %d %u %u\n", number, (int) wfa->frame_type,
		 wfa->states, wfa->root_state);
	 for (state = wfa->basis_states; state < wfa->states; state++)
	 {
	    printf ("s %u %d %d %d %u %u %u %u %d %d\n", state,
		    (int) wfa->level_of_state [state],
		    (int) wfa->tree [state][0], (int) wfa->tree [state][1],
		    (unsigned) wfa->x [state][0], (unsigned) wfa->y [state][0],
		    (unsigned) wfa->x [state][1], (unsigned) wfa->y [state][1],
		    (int) wfa->prediction [state][0],
		    (int) wfa->prediction [state][1]);
	    /* motion vectors and delta flags (predicted frames only: the lines are absent
	       from intra frames, so the still-image goldens do not change) */
	    for (label = 0; label < MAXLABELS; label++)
	       if (wfa->mv_tree [state][label].type != NONE)
		  printf ("m %u %u %d %d %d %d %d\n", state, label,
			  (int) wfa->mv_tree [state][label].type,
			  wfa->mv_tree [state][label].fx, wfa->mv_tree [state][label].fy,
			  wfa->mv_tree [state][label].bx, wfa->mv_tree [state][label].by);
	    if (wfa->delta_state [state])
	       printf ("d %u\n", state);
	    for (label = 0; label < MAXLABELS; label++)
	       for (edge = 0; isedge (wfa->into [state][label][edge]); edge++)
		  printf ("e %u %u %d %08x %.9g\n", state, label,
			  (int) wfa->into [state][label][edge],
			  float_bits (wfa->weight [state][label][edge]),
			  (double) wfa->weight [state][label][edge]);
	 }
	 printf ("end\n");
	 remove_states (wfa->basis_states, wfa);
      }
      close_bitfile (input);
      return 0;
   }
   catch
   {
      fprintf (stderr, "wfadump: %s\n", fiasco_get_error_message ());
      return 1;
   }
}
