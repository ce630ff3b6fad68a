/*
 *  fco_writer.c -- serialises a finished automaton into the FIASCO bit stream (.fco).
 *
 *  The stream layout and every coding decision must equal the reference's writer so that
 *  the files are byte-identical and decodable by the reference dfiasco:
 *    header / frame header   output/write.c:53-213
 *    bintree		      output/tree.c:46-190      (breadth first, adaptive binary coder)
 *    matrices		      output/matrices.c:53-536  (DC column, #edges, index deltas, chroma)
 *    weights		      output/weights.c:38-200   (per-level adaptive array coder)
 *    nondeterminism	      output/nd.c:53-244        (streams coded with `--prediction')
 *    motion		      output/mc.c:75-251        (P and B frames)
 *  The tiling flag is emitted as 0 (the reference's tiling never takes effect).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fi_internal.h"

#define RICE_K	   8
#define MIN_PROB   1
#define MAX_PROB   9
#define isrange(x) ((x) == FI_RANGE)
#define usedomain(s, w) ((w)->domain_type [s] & FI_USE_DOMAIN)

void
fi_write_header (const fi_wfainfo_t *wi, fi_bits_t *out)
{
   const char *t;

   for (t = "FIASCO"; *t; t++)
      fi_put_bits (out, (unsigned char) *t, 8);
   fi_put_bits (out, '\n', 8);
   for (t = wi->basis_name; *t; t++)
      fi_put_bits (out, (unsigned char) *t, 8);
   fi_put_bits (out, 0, 8);

   fi_write_rice (out, 2, RICE_K);		/* FIASCO_BINFILE_RELEASE */
   fi_write_rice (out, 1, RICE_K);		/* HEADER_TITLE */
   for (t = wi->title; t && *t && t - wi->title < FI_MAXSTRLEN - 2; t++)
      fi_put_bits (out, (unsigned char) *t, 8);
   fi_put_bits (out, 0, 8);
   fi_write_rice (out, 2, RICE_K);		/* HEADER_COMMENT */
   for (t = wi->comment; t && *t && t - wi->comment < FI_MAXSTRLEN - 2; t++)
      fi_put_bits (out, (unsigned char) *t, 8);
   fi_put_bits (out, 0, 8);
   fi_write_rice (out, 0, RICE_K);		/* HEADER_END */

   fi_write_rice (out, wi->max_states, RICE_K);
   fi_put_bit (out, wi->color ? 1 : 0);
   fi_write_rice (out, wi->width, RICE_K);
   fi_write_rice (out, wi->height, RICE_K);
   if (wi->color)
      fi_write_rice (out, wi->chroma_max_states, RICE_K);
   fi_write_rice (out, wi->p_min_level, RICE_K);
   fi_write_rice (out, wi->p_max_level, RICE_K);
   fi_write_rice (out, wi->frames, RICE_K);
   fi_write_rice (out, wi->smoothing, RICE_K);

   fi_put_bits (out, wi->rpf.mantissa_bits - 2, 3);
   fi_put_bits (out, (unsigned) wi->rpf.range_e, 2);
   {
      const fi_rpf_t *pair [3][2] = {{&wi->rpf, &wi->dc_rpf},
				     {&wi->rpf, &wi->d_rpf},
				     {&wi->dc_rpf, &wi->d_dc_rpf}};
      int i;

      for (i = 0; i < 3; i++)
	 if (pair [i][0]->mantissa_bits != pair [i][1]->mantissa_bits
	     || pair [i][0]->range != pair [i][1]->range)
	 {
	    fi_put_bit (out, 1);
	    fi_put_bits (out, pair [i][1]->mantissa_bits - 2, 3);
	    fi_put_bits (out, (unsigned) pair [i][1]->range_e, 2);
	 }
	 else
	    fi_put_bit (out, 0);
   }
   if (wi->frames > 1)
   {
      fi_write_rice (out, wi->fps, RICE_K);
      fi_write_rice (out, wi->search_range, RICE_K);
      fi_put_bit (out, wi->half_pixel ? 1 : 0);
      fi_put_bit (out, wi->B_as_past_ref ? 1 : 0);
   }
   fi_byte_align (out);
}

/* ---------------------------------------------------------------- bintree ---- */

static void
write_tree (const fi_wfa_t *wfa, fi_bits_t *out)
{
   unsigned *queue = fiasco_calloc (FI_MAXSTATES, sizeof (unsigned));
   uint8_t  *bits  = fiasco_calloc (FI_MAXSTATES * FI_MAXLABELS, 1);
   unsigned  last = 1, current, label, total = 0, n;
   fi_ac_t   ac;
   uint16_t  sum0 = 1, sum1 = 11;
   unsigned  scaling;

   /* breadth first order from the root: 1 = inner node, 0 = range */
   queue [0] = wfa->root_state;
   for (current = 0; current < last; current++)
      for (label = 0; label < FI_MAXLABELS; label++)
      {
	 const int into = wfa->tree [queue [current]][label];

	 if (!isrange (into))
	 {
	    queue [last++] = (unsigned) into;
	    bits [total++] = 1;
	 }
	 else
	    bits [total++] = 0;
      }
   if (total != (wfa->states - wfa->basis_states) * FI_MAXLABELS)
      fi_error ("total [%d] != (states - basis_states) * 2 [%d]", total,
		(wfa->states - wfa->basis_states) * FI_MAXLABELS);

   /* adaptive binary interval coder, counts (1, 11), halved beyond total / 20 */
   scaling = total / 20;
   fi_ac_init (&ac, out);
   for (n = 0; n < total; n++)
   {
      const unsigned range = (unsigned) (ac.high - ac.low) + 1;

      if (!bits [n])
      {
	 ac.high = (uint16_t) (ac.low + (uint16_t) ((range * sum0) / sum1 - 1));
	 fi_ac_rescale (&ac);
	 sum0++;
      }
      else
      {
	 ac.low = (uint16_t) (ac.low + (uint16_t) ((range * sum0) / sum1));
	 fi_ac_rescale (&ac);
      }
      sum1++;
      if (sum1 > scaling)
      {
	 sum0 >>= 1;
	 sum1 >>= 1;
	 if (!sum0)
	    sum0 = 1;
	 if (sum0 >= sum1)
	    sum1 = (uint16_t) (sum0 + 1);
      }
   }
   fi_ac_flush (&ac);
   free (queue);
   free (bits);
}

/* --------------------------------------------------------------- matrices ---- */

/* quasi arithmetic coder state used for the sparse 0/1 columns: the probability of the
   '1' symbol is 2^-prob[index], index walks up on '0' and is halved on '1' */
typedef struct qac
{
   fi_ac_t  ac;
   unsigned index;
} qac_t;

static unsigned prob_table [1 << (MAX_PROB + 1)];

static void
init_prob_table (void)
{
   unsigned index = 0, n, e;

   for (n = MIN_PROB; n <= MAX_PROB; n++)
      for (e = 0; e < 1u << n; e++, index++)
	 prob_table [index] = n;
}

static void
qac_encode (qac_t *q, int one)
{
   fi_ac_t *ac = &q->ac;

   if (!one)
   {
      ac->high = (uint16_t) (ac->high - ((ac->high - ac->low) >> prob_table [q->index]) - 1);
      fi_ac_rescale (ac);
      if (q->index < 1020)
	 q->index++;
   }
   else
   {
      ac->low = (uint16_t) (ac->high - ((ac->high - ac->low) >> prob_table [q->index]));
      fi_ac_rescale (ac);
      q->index >>= 1;
   }
}

static unsigned
n_edges (const fi_wfa_t *wfa, unsigned state, unsigned label)
{
   unsigned e = 0;

   while (wfa->into [state][label][e] != FI_NO_EDGE)
      e++;
   return e;
}

/* which ranges use the DC domain (state 0): one QAC bit per range of the luminance part */
static unsigned
column_0_encoding (const fi_wfa_t *wfa, unsigned last_row, fi_bits_t *out)
{
   qac_t    q;
   unsigned row, label, total = 0;

   fi_ac_init (&q.ac, out);
   q.index = 0;
   for (row = wfa->basis_states; row <= last_row; row++)
      for (label = 0; label < FI_MAXLABELS; label++)
	 if (isrange (wfa->tree [row][label]))
	 {
	    const int one = wfa->into [row][label][0] == 0;

	    qac_encode (&q, one);
	    total += (unsigned) one;
	 }
   fi_ac_flush (&q.ac);
   return total;
}

/* ranges in coder order with the largest domain each of them could refer to
   (codec/wfalib.c:659-696, including the slot reuse for a subdivided label 0) */
typedef struct range_list
{
   uint16_t *state, *max_domain;
   uint8_t  *label, *subdivided;
   unsigned  n;
} range_list_t;

static void
sort_ranges (unsigned state, unsigned *domain, range_list_t *rs, const fi_wfa_t *wfa)
{
   unsigned label;

   for (label = 0; label < FI_MAXLABELS; label++)
   {
      if (isrange (wfa->tree [state][label]))
	 rs->subdivided [rs->n] = 0;
      else
      {
	 sort_ranges ((unsigned) wfa->tree [state][label], domain, rs, wfa);
	 rs->subdivided [rs->n] = 1;
      }
      rs->state [rs->n]	     = (uint16_t) state;
      rs->label [rs->n]	     = (uint8_t) label;
      rs->max_domain [rs->n] = (uint16_t) *domain;
      while (!usedomain (rs->max_domain [rs->n], wfa))
	 rs->max_domain [rs->n]--;
      if (label == 1 || !rs->subdivided [rs->n])
	 rs->n++;
   }
   (*domain)++;
}

static unsigned
delta_encoding (int use_normal_domains, int use_delta_domains, const fi_wfa_t *wfa,
		unsigned last_domain, fi_bits_t *out)
{
   range_list_t rs;
   unsigned	max_domain, total = 0;
   unsigned	count [FI_MAXEDGES + 1];
   unsigned	state, label, n, M = 0, range;
   const size_t slots = (size_t) (last_domain + 1) * FI_MAXLABELS;

   rs.state	 = fiasco_calloc (slots, sizeof (uint16_t));
   rs.max_domain = fiasco_calloc (slots, sizeof (uint16_t));
   rs.label	 = fiasco_calloc (slots, 1);
   rs.subdivided = fiasco_calloc (slots, 1);
   rs.n		 = 0;
   max_domain	 = wfa->basis_states - 1;
   sort_ranges (last_domain, &max_domain, &rs, wfa);

   /* distribution of the number of edges per range, then the numbers themselves with a
      static model built from that distribution */
   for (n = 0; n < FI_MAXEDGES + 1; n++)
      count [n] = 0;
   for (state = wfa->basis_states; state <= last_domain; state++)
      for (label = 0; label < FI_MAXLABELS; label++)
	 if (isrange (wfa->tree [state][label]))
	 {
	    const unsigned e = n_edges (wfa, state, label);

	    count [e]++;
	    if (e > M)
	       M = e;
	 }
   fi_write_rice (out, M, 3);
   for (n = 0; n <= M; n++)
      fi_write_rice (out, count [n], (unsigned) ((int) log2 ((double) last_domain) - 2));
   {
      /* static order-0 model: cumulative counts, symbol s owns [cum[s], cum[s+1]) */
      unsigned cum [FI_MAXEDGES + 2];
      fi_ac_t  ac;

      cum [0] = 0;
      for (n = 0; n <= M; n++)
	 cum [n + 1] = cum [n] + count [n];
      fi_ac_init (&ac, out);
      for (range = 0; range < rs.n; range++)
	 if (!rs.subdivided [range])
	 {
	    const unsigned e	 = n_edges (wfa, rs.state [range], rs.label [range]);
	    const unsigned width = (unsigned) (ac.high - ac.low) + 1;
	    const uint16_t scale = (uint16_t) cum [M + 1];
	    const uint16_t lo	 = (uint16_t) cum [e], hi = (uint16_t) cum [e + 1];

	    ac.high = (uint16_t) (ac.low + (uint16_t) ((width * hi) / scale - 1));
	    ac.low  = (uint16_t) (ac.low + (uint16_t) ((width * lo) / scale));
	    fi_ac_rescale (&ac);
	 }
      fi_ac_flush (&ac);
   }

   /* the domain indices: ascending per range, coded as differences with an adjusted
      binary code whose alphabet shrinks to what is still possible */
   {
      uint16_t *mapping1 = fiasco_calloc (wfa->states, sizeof (uint16_t));
      uint16_t *mapping2 = fiasco_calloc (wfa->states, sizeof (uint16_t));
      unsigned	n1 = 0, n2 = 0;

      fi_put_bit (out, (unsigned) use_normal_domains);
      fi_put_bit (out, (unsigned) use_delta_domains);
      /* no delta states on an intra frame without prediction: both mappings count the
	 usable domains below each state */
      for (state = 0; state < wfa->states; state++)
      {
	 mapping1 [state] = (uint16_t) n1;
	 if (usedomain (state, wfa))
	    n1++;
	 mapping2 [state] = (uint16_t) n2;
	 if (usedomain (state, wfa)
	     && (state < wfa->basis_states || use_normal_domains))
	    n2++;
      }
      for (range = 0; range < rs.n; range++)
	 if (!rs.subdivided [range])
	 {
	    const unsigned st	     = rs.state [range], lb = rs.label [range];
	    const unsigned max_value = mapping1 [rs.max_domain [range]];
	    unsigned	   last	     = 1, edge;
	    int		   domain;

	    for (edge = 0; (domain = wfa->into [st][lb][edge]) != FI_NO_EDGE; edge++)
	       if (domain > 0)
	       {
		  total++;
		  if (max_value - last)
		  {
		     fi_write_bin_code (out, mapping1 [domain] - last, max_value - last);
		     last = mapping1 [domain] + 1u;
		  }
	       }
	 }
      free (mapping1);
      free (mapping2);
   }
   free (rs.state);
   free (rs.max_domain);
   free (rs.label);
   free (rs.subdivided);
   return total;
}

static int
cmp_hits (const void *a, const void *b)
{
   /* descending by hit count (lib/misc.c sort_desc_pair) */
   return (int) ((const int16_t *) b) [0] - (int) ((const int16_t *) a) [0];
}

static int
cmp_word (const void *a, const void *b)
{
   return (int) *(const int16_t *) a - (int) *(const int16_t *) b;
}

/* the n most referenced states among the luminance states (codec/wfalib.c:182-231) */
static int16_t *
compute_hits (unsigned from, unsigned to, unsigned n, const fi_wfa_t *wfa)
{
   int16_t (*hits) [2] = fiasco_calloc (to, sizeof *hits);	/* {key, value} */
   int16_t *domains;
   unsigned state, label, edge;
   int	    domain;

   for (domain = 0; domain < (int) to; domain++)
   {
      hits [domain][0] = 0;
      hits [domain][1] = (int16_t) domain;
   }
   for (state = from; state <= to; state++)
      for (label = 0; label < FI_MAXLABELS; label++)
	 for (edge = 0; (domain = wfa->into [state][label][edge]) != FI_NO_EDGE; edge++)
	    hits [domain][0]++;
   qsort (hits + 1, to - 1, sizeof *hits, cmp_hits);
   if (n > to)
      n = to;
   domains = fiasco_calloc (n + 1, sizeof (int16_t));
   for (domain = 0; domain < (int) n && (!domain || hits [domain][0]); domain++)
      domains [domain] = hits [domain][1];
   n = (unsigned) domain;
   qsort (domains, n, sizeof (int16_t), cmp_word);
   domains [n] = -1;
   free (hits);
   return domains;
}

static unsigned
chroma_encoding (const fi_wfa_t *wfa, fi_bits_t *out)
{
   const unsigned y_root = (unsigned) wfa->tree [wfa->tree [wfa->root_state][0]][0];
   int16_t	 *y_domains = compute_hits (wfa->basis_states, y_root,
					    wfa->info->chroma_max_states, wfa);
   qac_t	  q;
   unsigned	  domain, row, label, total = 0, next_index = 0;

   fi_ac_init (&q.ac, out);
   q.index = 0;
   /* one column per admitted domain; the probability index restarts for every column
      at the value it had after the first row of the previous column */
   for (domain = 0; y_domains [domain] != -1; domain++)
   {
      int save_index = 1;

      q.index = next_index;
      for (row = y_root + 1; row < wfa->states; row++)
      {
	 for (label = 0; label < FI_MAXLABELS; label++)
	    if (isrange (wfa->tree [row][label]))
	    {
	       unsigned edge;
	       int	into, match = 0;

	       for (edge = 0; (into = wfa->into [row][label][edge]) != FI_NO_EDGE
			      && (unsigned) into < row; edge++)
		  if (into == y_domains [domain] && into != wfa->y_state [row][label])
		     match = 1;
	       qac_encode (&q, match);
	       total += (unsigned) match;
	    }
	 if (save_index)
	 {
	    next_index = q.index;
	    save_index = 0;
	 }
      }
   }
   /* the extra column: does the range refer to the state at the same position in Y? */
   q.index = 0;
   for (row = y_root + 1; row < wfa->states; row++)
      for (label = 0; label < FI_MAXLABELS; label++)
      {
	 const int one = wfa->y_column [row][label] != 0;

	 qac_encode (&q, one);
	 total += (unsigned) one;
      }
   fi_ac_flush (&q.ac);
   free (y_domains);
   return total;
}

static unsigned
write_matrices (int use_normal_domains, int use_delta_domains, const fi_wfa_t *wfa,
		fi_bits_t *out)
{
   const unsigned root = wfa->info->color
			 ? (unsigned) wfa->tree [wfa->tree [wfa->root_state][0]][0]
			 : wfa->root_state;
   unsigned total;

   total  = column_0_encoding (wfa, root, out);
   total += delta_encoding (use_normal_domains, use_delta_domains, wfa, root, out);
   if (wfa->info->color)
      total += chroma_encoding (wfa, out);
   return total;
}

/* ---------------------------------------------------------------- weights ---- */

static void
write_weights (unsigned total, const fi_wfa_t *wfa, fi_bits_t *out)
{
   unsigned  state, label, offset1, offset2, offset3, offset4, i;
   unsigned *weights = fiasco_calloc (total, sizeof (unsigned));
   unsigned *levels  = fiasco_calloc (total, sizeof (unsigned));
   unsigned *c_symbols;
   unsigned  n = 0;
   int	     min_level = FI_MAXLEVEL, max_level = 0, dc = 0;
   int	     d_min_level = FI_MAXLEVEL, d_max_level = 0, d_dc = 0, delta_approx = 0;

   /* has a delta approximation (prediction error of a predicted range) been used?
      (output/weights.c:60-68) */
   if (wfa->delta_state)
      for (state = wfa->basis_states; state < wfa->states; state++)
	 if (wfa->delta_state [state])
	 {
	    delta_approx = 1;
	    break;
	 }
#define IS_DELTA(s) (delta_approx && wfa->delta_state [s])
   for (state = wfa->basis_states; state < wfa->states; state++)
      for (label = 0; label < FI_MAXLABELS; label++)
	 if (isrange (wfa->tree [state][label]))
	 {
	    const int l = (int) wfa->level_of_state [state] - 1;

	    if (IS_DELTA (state))
	    {
	       if (l < d_min_level)
		  d_min_level = l;
	       if (l > d_max_level)
		  d_max_level = l;
	       if (wfa->into [state][label][0] == 0)
		  d_dc = 1;
	    }
	    else
	    {
	       if (l < min_level)
		  min_level = l;
	       if (l > max_level)
		  max_level = l;
	       if (wfa->into [state][label][0] == 0)
		  dc = 1;
	    }
	 }
   if (min_level > max_level)
      max_level = min_level - 1;
   if (d_min_level > d_max_level)
      d_max_level = d_min_level - 1;
   /* contexts (output/weights.c:104-115): [0] DC weights, [1] delta DC weights, one per
      range level, one per delta range level */
   offset1 = dc ? 1 : 0;
   offset2 = offset1 + (d_dc ? 1 : 0);
   offset3 = offset2 + (unsigned) (max_level - min_level + 1);
   offset4 = offset3 + (unsigned) (d_max_level - d_min_level + 1);

   for (state = wfa->basis_states; state < wfa->states; state++)
      for (label = 0; label < FI_MAXLABELS; label++)
	 if (isrange (wfa->tree [state][label]))
	 {
	    unsigned edge;
	    int	     domain;

	    for (edge = 0; (domain = wfa->into [state][label][edge]) != FI_NO_EDGE; edge++)
	    {
	       if (n >= total)
		  fi_error ("Can't write more than %d weights.", total);
	       if (domain)
	       {
		  if (IS_DELTA (state))
		  {
		     weights [n] = (unsigned) fi_rtob (wfa->weight [state][label][edge],
						       &wfa->info->d_rpf);
		     levels [n]	 = offset3 + (unsigned) ((int) wfa->level_of_state [state]
							 - 1 - d_min_level);
		  }
		  else
		  {
		     weights [n] = (unsigned) fi_rtob (wfa->weight [state][label][edge],
						       &wfa->info->rpf);
		     levels [n]	 = offset2 + (unsigned) ((int) wfa->level_of_state [state]
							 - 1 - min_level);
		  }
	       }
	       else
	       {
		  if (IS_DELTA (state))
		  {
		     weights [n] = (unsigned) fi_rtob (wfa->weight [state][label][edge],
						       &wfa->info->d_dc_rpf);
		     levels [n]	 = offset1;
		  }
		  else
		  {
		     weights [n] = (unsigned) fi_rtob (wfa->weight [state][label][edge],
						       &wfa->info->dc_rpf);
		     levels [n]	 = 0;
		  }
	       }
	       n++;
	    }
	 }
#undef IS_DELTA
   c_symbols	 = fiasco_calloc (offset4 ? offset4 : 1, sizeof (unsigned));
   c_symbols [0] = 1u << (wfa->info->dc_rpf.mantissa_bits + 1);
   if (offset1 != offset2)
      c_symbols [offset1] = 1u << (wfa->info->d_dc_rpf.mantissa_bits + 1);
   for (i = offset2; i < offset3; i++)
      c_symbols [i] = 1u << (wfa->info->rpf.mantissa_bits + 1);
   for (; i < offset4; i++)
      c_symbols [i] = 1u << (wfa->info->d_rpf.mantissa_bits + 1);
   fi_encode_array (out, weights, levels, c_symbols, offset4, total, 500);
   free (c_symbols);
   free (weights);
   free (levels);
}

/* ------------------------------------------------------- motion (output/mc.c) ---- */

/* MPEG's Huffman code of a vector component: value, length (codec/mwfa.c:40-52) */
static const unsigned short mv_code_table [33][2] =
{
   {0x19, 11}, {0x1b, 11}, {0x1d, 11}, {0x1f, 11}, {0x21, 11}, {0x23, 11}, {0x13, 10},
   {0x15, 10}, {0x17, 10}, {0x7, 8}, {0x9, 8}, {0xb, 8}, {0x7, 7}, {0x3, 5}, {0x3, 4},
   {0x3, 3}, {0x1, 1}, {0x2, 3}, {0x2, 4}, {0x2, 5}, {0x6, 7}, {0xa, 8}, {0x8, 8},
   {0x6, 8}, {0x16, 10}, {0x14, 10}, {0x12, 10}, {0x22, 11}, {0x20, 11}, {0x1e, 11},
   {0x1c, 11}, {0x1a, 11}, {0x18, 11}
};

/*
 *  write_mc (output/mc.c:75-84) for a predicted frame: the tree of motion compensation
 *  decisions in breadth-first order from the highest prediction level down (encode_mc_tree,
 *  :91-150; P frames: one bit per decision, 1 = none; B frames: none 1, forward 000,
 *  backward 001, interpolated 01), then the vector components in state order
 *  (encode_mc_coords, :152-251).
 */
static void
write_mc (const fi_wfa_t *wfa, fi_bits_t *out)
{
   const fi_wfainfo_t *wi	 = wfa->info;
   const unsigned      max_state = wi->color ? (unsigned) wfa->tree [wfa->tree [wfa->root_state][0]][0]
					     : wfa->states;
   unsigned	      *queue	 = fiasco_calloc (wfa->states + 1, sizeof (unsigned));
   unsigned	       last = 0, current, state, label;

   if (!wfa->mv_type || !wfa->x || !wfa->y)
      fi_error ("predicted frame without motion data");
   for (state = wfa->basis_states; state < max_state; state++)
      if ((int) wfa->level_of_state [state] - 1 == (int) wi->p_max_level)
	 queue [last++] = state;
   for (current = 0; current < last; current++)
      for (label = 0; label < FI_MAXLABELS; label++)
      {
	 state = queue [current];
	 const int	type  = wfa->mv_type [state][label];
	 const unsigned level = (unsigned) wfa->level_of_state [state] - 1;

	 if (wfa->x [state][label] + (1u << (level >> 1)) <= wi->width
	     && wfa->y [state][label] + (1u << ((level + 1) >> 1)) <= wi->height)
	 {
	    if (wfa->frame_type == 1)
	       fi_put_bit (out, type == 0);	/* p_frame_codes: none = 1, forward = 0 */
	    else			/* b_frame_codes (mc.c:46-53) */
	       fi_put_bits (out, type == 0 ? 1 : type == 3 ? 1 : type == 2 ? 1 : 0,
			    type == 0 ? 1 : type == 3 ? 2 : 3);
	 }
	 if (type == 0 && !isrange (wfa->tree [state][label])
	     && (int) level >= (int) wi->p_min_level)
	    queue [last++] = (unsigned) wfa->tree [state][label];
      }
   fi_byte_align (out);
   for (state = wfa->basis_states; state < max_state; state++)
      for (label = 0; label < FI_MAXLABELS; label++)
      {
	 const int type = wfa->mv_type [state][label];

	 if (type == 1 || type == 3)			/* forward vector */
	 {
	    const unsigned ix = (unsigned) (wfa->mv_fx [state][label] + (int) wi->search_range);
	    const unsigned iy = (unsigned) (wfa->mv_fy [state][label] + (int) wi->search_range);

	    fi_put_bits (out, mv_code_table [ix][0], mv_code_table [ix][1]);
	    fi_put_bits (out, mv_code_table [iy][0], mv_code_table [iy][1]);
	 }
	 if (type == 2 || type == 3)			/* backward vector */
	 {
	    if (!wfa->mv_bx || !wfa->mv_by)
	       fi_error ("B frame without backward vectors");
	    const unsigned ix = (unsigned) (wfa->mv_bx [state][label] + (int) wi->search_range);
	    const unsigned iy = (unsigned) (wfa->mv_by [state][label] + (int) wi->search_range);

	    fi_put_bits (out, mv_code_table [ix][0], mv_code_table [ix][1]);
	    fi_put_bits (out, mv_code_table [iy][0], mv_code_table [iy][1]);
	 }
      }
   fi_byte_align (out);
   free (queue);
}

/* ------------------------------------------------------------------ frame ---- */

/* ------------------------------------- nondeterministic prediction (output/nd.c) ---- */

/*
 *  A stream coded with `--prediction' carries, per frame, which ranges of the prediction levels
 *  are predicted by their DC component -- the subdivided labels that have edges as well, one
 *  adaptive bit each in breadth-first order (encode_nd_tree, output/nd.c:70-190: counts (1, 11),
 *  halved beyond 50; below a predicted label the walk stops) -- and the weights of those edges
 *  (encode_nd_coefficients, :192-244: one uniform-start table in the DC format).
 */
static void
write_nd (const fi_wfa_t *wfa, fi_bits_t *out)
{
   unsigned *queue = fiasco_calloc (FI_MAXSTATES, sizeof (unsigned));
   unsigned  head = 0, tail = 0, used = 0;
   fi_ac_t   ac;
   uint16_t  sum0 = 1, sum1 = 11;

   fi_ac_init (&ac, out);
   queue [tail++] = wfa->root_state;
   while (head < tail)
   {
      const unsigned next  = queue [head++];
      const unsigned level = wfa->level_of_state [next];

      if (level > wfa->info->p_max_level + 1)
      {
	 for (unsigned label = 0; label < FI_MAXLABELS; label++)
	    if (!isrange (wfa->tree [next][label]))
	       queue [tail++] = (unsigned) wfa->tree [next][label];
      }
      else if (level > wfa->info->p_min_level)
	 for (unsigned label = 0; label < FI_MAXLABELS; label++)
	 {
	    const int child = wfa->tree [next][label];

	    if (isrange (child))
	       continue;
	    const unsigned range = (unsigned) (ac.high - ac.low) + 1;

	    if (wfa->into [next][label][0] != FI_NO_EDGE)	/* predicted */
	    {
	       used++;
	       ac.low = (uint16_t) (ac.low + (uint16_t) ((range * sum0) / sum1));
	       fi_ac_rescale (&ac);
	    }
	    else
	    {
	       if (wfa->level_of_state [child] > wfa->info->p_min_level)
		  queue [tail++] = (unsigned) child;
	       ac.high = (uint16_t) (ac.low + (uint16_t) ((range * sum0) / sum1 - 1));
	       fi_ac_rescale (&ac);
	       sum0++;
	    }
	    sum1++;
	    if (sum1 > 50)
	    {
	       sum0 >>= 1;
	       sum1 >>= 1;
	       if (!sum0)
		  sum0 = 1;
	       if (sum0 >= sum1)
		  sum1 = (uint16_t) (sum0 + 1);
	    }
	 }
   }
   fi_ac_flush (&ac);
   free (queue);

   if (used)
   {
      unsigned *coeff = fiasco_calloc (used * FI_MAXEDGES, sizeof (unsigned));
      unsigned	n = 0, c_symbols = 1u << (wfa->info->dc_rpf.mantissa_bits + 1);

      for (unsigned state = wfa->basis_states; state < wfa->states; state++)
	 for (unsigned label = 0; label < FI_MAXLABELS; label++)
	    if (!isrange (wfa->tree [state][label]))
	       for (unsigned edge = 0; wfa->into [state][label][edge] != FI_NO_EDGE; edge++)
	       {
		  if (n >= used)
		     fi_error ("Can't write more than %d coefficients.", used);
		  coeff [n++] = (unsigned) fi_rtob (wfa->weight [state][label][edge], &wfa->info->dc_rpf);
	       }
      fi_encode_array (out, coeff, NULL, &c_symbols, 1, used, 50);
      free (coeff);
   }
}

/* the writer's static tables, before host threads share them */
void
fi_write_tables_init (void)
{
   if (!prob_table [0])
      init_prob_table ();
}

void
fi_write_next_wfa (const fi_wfa_t *wfa, unsigned frame_number, int first,
		   int normal_domains, int delta_domains, fi_bits_t *out)
{
   unsigned edges;

   if (!prob_table [0])
      init_prob_table ();
   if (first)
      fi_write_header (wfa->info, out);
   fi_write_rice (out, wfa->states, RICE_K);
   fi_write_rice (out, (unsigned) wfa->frame_type, RICE_K);	/* 0 I_FRAME, 1 P_FRAME */
   fi_write_rice (out, frame_number, RICE_K);
   fi_byte_align (out);
   fi_put_bit (out, 0);				/* no tiling permutation */
   fi_byte_align (out);
   write_tree (wfa, out);
   if (wfa->info->nd_prediction)			/* output/write.c:100-106 */
   {
      fi_put_bit (out, 1);
      write_nd (wfa, out);
   }
   else
      fi_put_bit (out, 0);
   if (wfa->frame_type != 0)
      write_mc (wfa, out);
   edges = write_matrices (normal_domains, delta_domains, wfa, out);
   if (edges)
      write_weights (edges, wfa, out);
}
