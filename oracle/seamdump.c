/*
 *  seamdump.c -- TEST INFRASTRUCTURE (oracle side), not part of the product.
 *
 *  Runs the UNMODIFIED reference coder (fiasco_coder(), linked from
 *  oracle/_ref/libfiasco_ref.a) on one input and records what crosses the reference's
 *  inner seams of the hot path, using GNU ld's --wrap (no reference source is patched
 *  or copied):
 *
 *    approximate_range()        codec/approx.c:74      -> one "lc" line per call
 *    append_state()             codec/control.c:48     -> one "st" line per call
 *    compute_ip_images_state()  codec/ip.c:72          -> "ipis" lines (range x state rows)
 *                                                         for the first blocks (level == lc_max)
 *
 *  It can also print known-answer tables for the small pure functions on the path
 *  (rtob/btor lib/rpf.c:59,113; bits_bin_code lib/misc.c:296; tree_bits
 *  codec/bintree.c:55) with `seamdump --kat`.
 *
 *  usage: seamdump [--kat] | seamdump <in.pnm> <out.fco> <quality> <optimize 0..3> [trace.txt]
 *         (CLI option mapping follows bin/cwfa.c:326-345: -z0 => levels [6,10], 3 edges)
 *
 *  Link line (see oracle/Makefile): -Wl,--wrap=approximate_range,--wrap=append_state,
 *  --wrap=compute_ip_images_state
 */
#include "config.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "types.h"
#include "macros.h"
#include "error.h"
#include "cwfa.h"
#include "approx.h"
#include "control.h"
#include "ip.h"
#include "rpf.h"
#include "misc.h"
#include "bintree.h"
#include "mwfa.h"
#include "prediction.h"
#include "fiasco.h"

static FILE    *trace      = NULL;
static unsigned lc_calls   = 0;
static unsigned ipis_calls = 0;

static unsigned
fbits (float f)
{
   unsigned u;
   memcpy (&u, &f, sizeof u);
   return u;
}

real_t
__real_approximate_range (real_t max_costs, real_t price, int max_edges,
			  int y_state, range_t *range,
			  domain_pool_t *domain_pool, coeff_t *coeff,
			  const wfa_t *wfa, const coding_t *c);
void
__real_append_state (bool_t auxiliary_state, real_t final,
		     unsigned level_of_state, wfa_t *wfa, coding_t *c);
void
__real_compute_ip_images_state (unsigned image, unsigned address,
				unsigned level, unsigned n, unsigned from,
				const wfa_t *wfa, coding_t *c);

/* predicted frames: the motion vector every mc_prediction() finds and the outcome of every
   predict_range() (codec/mwfa.c:301, codec/prediction.c:96) */
void
__real_find_P_frame_mc (word_t *mcpe, real_t price, range_t *range,
			const wfa_info_t *wi, const motion_t *mt);
real_t
__real_predict_range (real_t max_costs, real_t price, range_t *range, wfa_t *wfa,
		      coding_t *c, unsigned band, int y_state, unsigned states,
		      const tree_t *tree_model, const tree_t *p_tree_model,
		      const void *domain_model, const void *d_domain_model,
		      const void *coeff_model, const void *d_coeff_model);

void
__wrap_find_P_frame_mc (word_t *mcpe, real_t price, range_t *range,
			const wfa_info_t *wi, const motion_t *mt)
{
   __real_find_P_frame_mc (mcpe, price, range, wi, mt);
   if (trace)
      fprintf (trace, "mv %u %u %u %d %d %08x\n", range->level, range->x, range->y,
	       range->mv.fx, range->mv.fy, fbits (range->mv_coord_bits));
}

real_t
__wrap_predict_range (real_t max_costs, real_t price, range_t *range, wfa_t *wfa,
		      coding_t *c, unsigned band, int y_state, unsigned states,
		      const tree_t *tree_model, const tree_t *p_tree_model,
		      const void *domain_model, const void *d_domain_model,
		      const void *coeff_model, const void *d_coeff_model)
{
   const unsigned level = range->level, x = range->x, y = range->y;
   const unsigned before = wfa->states;
   real_t	  costs;

   costs = __real_predict_range (max_costs, price, range, wfa, c, band, y_state, states,
				 tree_model, p_tree_model, domain_model, d_domain_model,
				 coeff_model, d_coeff_model);
   if (trace)
      fprintf (trace, "pr %u %u %u %u %u %08x %08x %u\n", level, x, y, states, before,
	       fbits (max_costs), fbits (costs), wfa->states);
   return costs;
}

real_t
__wrap_approximate_range (real_t max_costs, real_t price, int max_edges,
			  int y_state, range_t *range,
			  domain_pool_t *domain_pool, coeff_t *coeff,
			  const wfa_t *wfa, const coding_t *c)
{
   real_t costs = __real_approximate_range (max_costs, price, max_edges,
					    y_state, range, domain_pool,
					    coeff, wfa, c);
   if (trace)
   {
      int e;

      fprintf (trace, "lc %u %u %u %u %u %u %d %u %08x %08x %08x",
	       lc_calls, range->level, range->image, range->address,
	       range->x, range->y, y_state, wfa->states,
	       fbits (max_costs), fbits (price), fbits (costs));
      if (isedge (range->into [0]))
      {
	 fprintf (trace, " %08x %08x %08x :", fbits (range->err),
		  fbits (range->matrix_bits), fbits (range->weights_bits));
	 for (e = 0; isedge (range->into [e]); e++)
	    fprintf (trace, " %d:%08x", (int) range->into [e],
		     fbits (range->weight [e]));
      }
      fprintf (trace, "\n");
   }
   lc_calls++;
   return costs;
}

void
__wrap_append_state (bool_t auxiliary_state, real_t final,
		     unsigned level_of_state, wfa_t *wfa, coding_t *c)
{
   unsigned state = wfa->states;

   __real_append_state (auxiliary_state, final, level_of_state, wfa, c);
   if (trace)
   {
      fprintf (trace, "st %u %d %u %08x %d %d\n", state, (int) auxiliary_state,
	       level_of_state, fbits (final),
	       (int) wfa->tree [state][0], (int) wfa->tree [state][1]);
      if (!auxiliary_state && state < 12) /* a few known-answer tables */
      {
	 unsigned i, level, t;

	 fprintf (trace, "img %u", state);
	 for (i = 0; i < size_of_tree (c->options.images_level); i++)
	    fprintf (trace, " %08x", fbits (c->images_of_state [state][i]));
	 fprintf (trace, "\n");
	 for (level = c->options.images_level + 1;
	      level <= c->options.lc_max_level; level++)
	 {
	    fprintf (trace, "ipss %u %u", state, level);
	    for (t = 0; t <= state; t++)
	       fprintf (trace, " %08x",
			need_image (t, wfa)
			? fbits (c->ip_states_state [state][level][t]) : 0);
	    fprintf (trace, "\n");
	 }
      }
   }
}

void
__wrap_compute_ip_images_state (unsigned image, unsigned address,
				unsigned level, unsigned n, unsigned from,
				const wfa_t *wfa, coding_t *c)
{
   __real_compute_ip_images_state (image, address, level, n, from, wfa, c);
   /* the outermost call made by init_range() (codec/subdivide.c:643) */
   if (trace && image == 0 && from == 0 && n == 1
       && level == c->options.lc_max_level && ipis_calls < 4)
   {
      unsigned state, i;

      fprintf (trace, "pix %u", ipis_calls);
      for (i = 0; i < size_of_level (level); i++)
	 fprintf (trace, " %d", (int) c->pixels [i]);
      fprintf (trace, "\n");
      for (state = 0; state < wfa->states && state < 40; state++)
	 if (need_image (state, wfa))
	 {
	    fprintf (trace, "ipis %u %u", ipis_calls, state);
	    for (i = 0; i < size_of_tree (c->products_level); i++)
	       fprintf (trace, " %08x", fbits (c->ip_images_state [state][i]));
	    fprintf (trace, "\n");
	 }
      ipis_calls++;
   }
}

static void
known_answer_tables (void)
{
   static const fiasco_rpf_range_e ranges [] = {FIASCO_RPF_RANGE_0_75,
						FIASCO_RPF_RANGE_1_00,
						FIASCO_RPF_RANGE_1_50,
						FIASCO_RPF_RANGE_2_00};
   unsigned m, r, i;

   /* rtob / btor over a deterministic sweep of fp32 inputs */
   for (m = 2; m <= 8; m++)
      for (r = 0; r < 4; r++)
      {
	 rpf_t   *rpf = alloc_rpf (m, ranges [r]);
	 unsigned seed = 12345u + 97u * m + r;

	 for (i = 0; i < 1u << (m + 1); i++)
	    printf ("btor %u %u %u %08x\n", m, r, i, fbits (btor (i, rpf)));
	 for (i = 0; i < 600; i++)
	 {
	    float    f;
	    unsigned u;

	    seed = seed * 1664525u + 1013904223u;
	    if (i < 200)		/* dense sweep of [-2.5, 2.5] */
	       f = -2.5f + 0.025f * i;
	    else if (i < 400)		/* random magnitudes 2^-20 .. 2^6 */
	    {
	       u = ((107u + (seed >> 24) % 27u) << 23) | (seed & 0x807fffffu);
	       memcpy (&f, &u, 4);
	    }
	    else			/* small values around the zero threshold */
	       f = ((int) (seed >> 8) % 4001 - 2000) * 1e-4f;
	    printf ("rtob %u %u %08x %d\n", m, r, fbits (f), rtob (f, rpf));
	 }
      }
   /* bits_bin_code over all (value, maxval) with maxval <= 300 and a few big ones */
   for (m = 1; m <= 300; m++)
      for (i = 0; i <= m; i++)
	 printf ("bbc %u %u %u\n", i, m, bits_bin_code (i, m));
   for (m = 1000; m <= 6000; m += 617)
      for (i = 0; i <= m; i += 131)
	 printf ("bbc %u %u %u\n", i, m, bits_bin_code (i, m));
   /* tree model: initial tables and bits */
   {
      tree_t tree;

      init_tree_model (&tree);
      for (i = 0; i < MAXLEVEL; i++)
	 printf ("tree %u %u %u %08x %08x\n", i, tree.counts [i], tree.total [i],
		 fbits (tree_bits (CHILD, i, &tree)),
		 fbits (tree_bits (LEAF, i, &tree)));
   }
}

int
main (int argc, char **argv)
{
   if (argc == 2 && streq (argv [1], "--kat"))
   {
      known_answer_tables ();
      return 0;
   }
   if (argc < 5)
   {
      fprintf (stderr, "usage: %s --kat | %s in.pnm out.fco quality optimize "
	       "[trace.txt]\n", argv [0], argv [0]);
      return 2;
   }
   {
      const char	 *names [2] = {argv [1], NULL};
      float		  quality   = atof (argv [3]);
      int		  o	    = atoi (argv [4]);
      fiasco_c_options_t *options   = fiasco_c_options_new ();
      int		  M, m, N;

      if (argc > 5)
	 trace = fopen (argv [5], "w");
      fiasco_set_verbosity (FIASCO_NO_VERBOSITY);

      /* same calls, in the same order, as checkargs() of the reference CLI
	 (bin/cwfa.c:253-388) makes for its default parameter values */
      fiasco_c_options_set_frame_pattern (options, "ippppppppp");
      fiasco_c_options_set_basisfile (options, "small.fco");
      fiasco_c_options_set_chroma_quality (options, 2, 40);
      fiasco_c_options_set_smoothing (options, 70);
      fiasco_c_options_set_progress_meter (options, FIASCO_PROGRESS_NONE);
      fiasco_c_options_set_tiling (options, FIASCO_TILING_VARIANCE_DSC, 4);
      if (o <= 0)
      {
	 o = 0; M = 10; m = 6; N = 3;
      }
      else
      {
	 o -= 1; M = 12; m = 4; N = 5;
      }
      fiasco_c_options_set_optimizations (options, m, M, N, 10000, o);
      fiasco_c_options_set_prediction (options, 0, 6, 10);
      fiasco_c_options_set_quantization (options, 3, FIASCO_RPF_RANGE_1_50,
					 5, FIASCO_RPF_RANGE_1_00);
      if (!fiasco_coder (names, argv [2], quality, options))
      {
	 fprintf (stderr, "seamdump: %s\n", fiasco_get_error_message ());
	 return 1;
      }
      if (trace)
	 fclose (trace);
   }
   return 0;
}
