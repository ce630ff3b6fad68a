/*
 *  fiasco_oracle.c -- CPU restatement (in our own words) of the FIASCO encoder hot path:
 *  bintree subdivision, matching pursuit, inner-product tables, domain-pool and
 *  coefficient rate models.  Sequential, single-threaded, plain C.
 *
 *  TEST INFRASTRUCTURE ONLY (see fiasco_oracle.h).  Parity: PINNED against the
 *  reference binary built in oracle/_ref (tests/test_oracle_vs_reference.py and the
 *  committed golden dumps/traces in tests/golden/).
 *
 *  Every function cites the reference file:line whose behaviour it restates.  Floating
 *  point follows SURVEY.md Appendix A.7: real_t = float, every operation rounded to
 *  fp32, no contraction (built with -ffp-contract=off), double only inside log2().
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stddef.h>

#include "fiasco_oracle.h"

#define MAXEDGES  FO_MAXEDGES
#define MAXSTATES FO_MAXSTATES
#define MAXLABELS FO_MAXLABELS
#define MAXLEVEL  FO_MAXLEVEL
#define NO_EDGE	  (-1)
#define RANGE	  (-1)
#define AUXILIARY_MASK	1
#define USE_DOMAIN_MASK 2
#define MAXCOSTS  1e20f		/* codec/coder.c:53 */

#define width_of_level(l)   (1u << ((l) >> 1))		/* lib/macros.h:48 */
#define height_of_level(l)  (1u << (((l) + 1) >> 1))	/* lib/macros.h:49 */
#define size_of_level(l)    (1u << (l))
#define address_of_level(l) (size_of_level (l) - 1)
#define size_of_tree(l)	    (address_of_level ((l) + 1))
#define fmin2(a, b)	    ((a) > (b) ? (b) : (a))	/* macros.h:56 'min' */

/*****************************************************************************
			reduced precision format  (lib/rpf.c)
*****************************************************************************/

typedef struct rpf
{
   unsigned mantissa_bits;
   float    range;
} rpf_t;

static float
rpf_range_value (int range_e)	/* lib/rpf.c:202-222 */
{
   switch (range_e)
   {
      case 0:  return 0.75f;
      case 1:  return 1.00f;
      case 2:  return 1.50f;
      case 3:  return 2.00f;
      default: return 1.00f;
   }
}

static rpf_t
make_rpf (unsigned mantissa, int range_e) /* alloc_rpf, lib/rpf.c:171-199 */
{
   rpf_t r;

   if (mantissa < 2)
      mantissa = 2;
   else if (mantissa > 8)
      mantissa = 2;		/* sic: the reference "clamps" > 8 to 2 (quirk C4) */
   r.mantissa_bits = mantissa;
   r.range	   = rpf_range_value (range_e);
   return r;
}

static int
rtob (float f, const rpf_t *rpf)
/*
 *  lib/rpf.c:59-111.  The reference shifts a 32 bit word by 'exponent' which may
 *  exceed 31; compiled for x86-64 that is a 'shl/shr %cl' whose count is taken
 *  modulo 32.  We make that explicit so the restatement is well defined.
 */
{
   uint32_t u, mantissa;
   int	    exponent, sign;

   f /= rpf->range;
   memcpy (&u, &f, 4);
   mantissa = u & 0x7fffffu;
   exponent = (int) ((u >> 23) & 0xffu) - 126;
   sign	    = (int) (u >> 31);

   mantissa >>= 1;
   mantissa  |= 1u << 22;
   if (exponent > 0)
      mantissa <<= (exponent & 31);
   else
      mantissa >>= ((-exponent) & 31);
   mantissa >>= (23 - rpf->mantissa_bits - 1);
   mantissa  += 1;
   mantissa >>= 1;

   if (mantissa == 0)
      return -1;		/* RPF_ZERO */
   else if (mantissa >= (1u << rpf->mantissa_bits))
      return sign;
   else
      return (int) (((mantissa & ((1u << rpf->mantissa_bits) - 1)) << 1) | sign);
}

static float
btor (int binary, const rpf_t *rpf) /* lib/rpf.c:113-169 */
{
   uint32_t mantissa, u;
   int	    sign, exponent;
   float    v;

   if (binary == -1)
      return 0;
   sign	      = binary & 1;
   mantissa   = ((unsigned) binary & ((1u << (rpf->mantissa_bits + 1)) - 1)) >> 1;
   mantissa <<= (23 - rpf->mantissa_bits);
   exponent   = 0;
   if (mantissa == 0)
      v = sign ? -1.0f : 1.0f;
   else
   {
      while (!(mantissa & (1u << 22)))
      {
	 exponent--;
	 mantissa <<= 1;
      }
      mantissa <<= 1;
      u = ((uint32_t) sign << 31) | ((uint32_t) (exponent + 126) << 23)
	  | (mantissa & 0x7fffffu);
      memcpy (&v, &u, 4);
   }
   return v * rpf->range;
}

int
fo_rtob (float f, unsigned mantissa_bits, int range_e)
{
   rpf_t r = make_rpf (mantissa_bits, range_e);
   return rtob (f, &r);
}

float
fo_btor (int b, unsigned mantissa_bits, int range_e)
{
   rpf_t r = make_rpf (mantissa_bits, range_e);
   return btor (b, &r);
}

/*****************************************************************************
			     misc  (lib/misc.c)
*****************************************************************************/

static unsigned
bits_bin_code (unsigned value, unsigned maxval) /* lib/misc.c:296-315 */
{
   unsigned k = (unsigned) log2 ((double) (maxval + 1));
   unsigned r = (maxval + 1) % (1u << k);

   return value < maxval + 1 - 2 * r ? k : k + 1;
}

unsigned
fo_bits_bin_code (unsigned value, unsigned maxval)
{
   return bits_bin_code (value, maxval);
}

/* the form every rate term takes in the reference (bintree.c:64-67, coeff.c:231-236,
   domain-pool.c:774): -log2 (count / (real_t) total), double log2, narrowed to fp32 */
float
fo_neg_log2f (int count, int total)
{
   return (float) -log2 ((double) (count / (float) total));
}

/*****************************************************************************
			   tree model  (codec/bintree.c)
*****************************************************************************/

typedef struct tree_model
{
   unsigned counts [MAXLEVEL];
   unsigned total [MAXLEVEL];
} tree_model_t;

static void
init_tree_model (tree_model_t *t) /* codec/bintree.c:70-93 */
{
   static const unsigned c0 [MAXLEVEL] = {20, 17, 15, 10, 5, 4, 3, 2, 1, 1, 1,
					  1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
   static const unsigned c1 [MAXLEVEL] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 3, 5,
					  10, 15, 20, 25, 30, 35, 60, 60, 60,
					  60};
   unsigned l;

   for (l = 0; l < MAXLEVEL; l++)
   {
      t->counts [l] = c1 [l];
      t->total [l]  = c0 [l] + c1 [l];
   }
}

static float
tree_bits (int child, unsigned level, const tree_model_t *t) /* bintree.c:55-68 */
{
   float prob = t->counts [level] / (float) t->total [level];

   return child ? (float) -log2 ((double) prob)
		: (float) -log2 ((double) (1 - prob));
}

static void
tree_update (int child, unsigned level, tree_model_t *t) /* bintree.c:35-53 */
{
   if (child)
      t->counts [level]++;
   t->total [level]++;
}

void
fo_tree_model_kat (unsigned level, unsigned *counts, unsigned *total,
		   float *child_bits, float *leaf_bits)
{
   tree_model_t t;

   init_tree_model (&t);
   *counts     = t.counts [level];
   *total      = t.total [level];
   *child_bits = tree_bits (1, level, &t);
   *leaf_bits  = tree_bits (0, level, &t);
}

/*****************************************************************************
		 coefficient model "adaptive"  (codec/coeff.c:185-330)
*****************************************************************************/

#define AAC_MAXCOUNTS (1 + (1 << 9) + (MAXLEVEL + 1) * (1 << 9))

typedef struct aac_model
{
   /* counts_mem[0] is a pad slot: the reference can (in principle) index counts[-1]
      with the RPF_ZERO code; keep that read inside our own memory */
   int16_t counts_mem [AAC_MAXCOUNTS];
   int16_t totals [MAXLEVEL + 2];
} aac_model_t;

typedef struct coeff
{
   rpf_t    rpf, dc_rpf;
   unsigned min_level, max_level;
   aac_model_t model;
} coeff_t;

static void
aac_init (coeff_t *c, rpf_t rpf, rpf_t dc_rpf, unsigned min_level,
	  unsigned max_level) /* coeff.c:285-313 */
{
   unsigned size, n;

   c->rpf	= rpf;
   c->dc_rpf	= dc_rpf;
   c->min_level = min_level;
   c->max_level = max_level;
   size = (max_level - min_level + 1) * (1u << (1 + rpf.mantissa_bits))
	  + (1u << (1 + dc_rpf.mantissa_bits));
   memset (&c->model, 0, sizeof c->model);
   for (n = 0; n < size + 1; n++)
      c->model.counts_mem [n] = 1;
   c->model.totals [0] = (int16_t) (1 << (1 + dc_rpf.mantissa_bits));
   for (n = min_level; n <= max_level; n++)
      c->model.totals [n - min_level + 1]
	 = (int16_t) (1 << (1 + rpf.mantissa_bits));
}

static float
aac_bits (const float *used_coeff, const int16_t *used_states, unsigned level,
	  const coeff_t *c) /* coeff.c:215-240 */
{
   float	  bits = 0;
   unsigned	  edge;
   int		  state;
   const int16_t *base	 = c->model.counts_mem + 1;
   const int16_t *counts = base + (1 << (1 + c->dc_rpf.mantissa_bits))
			   + ((level - c->min_level)
			      * (1 << (1 + c->rpf.mantissa_bits)));

   for (edge = 0; (state = used_states [edge]) != NO_EDGE; edge++)
      if (state)
	 bits = (float) ((double) bits
			 - log2 ((double) (counts [rtob (used_coeff [edge], &c->rpf)]
					   / (float) c->model.totals [level - c->min_level + 1])));
      else
	 bits = (float) ((double) bits
			 - log2 ((double) (base [rtob (used_coeff [edge], &c->dc_rpf)]
					   / (float) c->model.totals [0])));
   return bits;
}

static void
aac_update (const float *used_coeff, const int16_t *used_states,
	    unsigned level, coeff_t *c) /* coeff.c:242-267 */
{
   unsigned edge;
   int	    state;
   int16_t *base   = c->model.counts_mem + 1;
   int16_t *counts = base + (1 << (1 + c->dc_rpf.mantissa_bits))
		     + ((level - c->min_level)
			* (1 << (1 + c->rpf.mantissa_bits)));

   for (edge = 0; (state = used_states [edge]) != NO_EDGE; edge++)
      if (state)
      {
	 counts [rtob (used_coeff [edge], &c->rpf)]++;
	 c->model.totals [level - c->min_level + 1]++;
      }
      else
      {
	 base [rtob (used_coeff [edge], &c->dc_rpf)]++;
	 c->model.totals [0]++;
      }
}

/*****************************************************************************
	      domain pool "rle" (+ its 1-entry "adaptive" DC model)
			(codec/domain-pool.c:621-879, :259-498, :970-999)
*****************************************************************************/

static float matrix_0 [1 << 10], matrix_1 [1 << 10];

static void
init_matrix_probabilities (void) /* domain-pool.c:970-999 */
{
   unsigned index = 0, n, e;

   for (n = 1; n <= 9; n++)
      for (e = 0; e < 1u << n; e++, index++)
      {
	 matrix_1 [index] = (float) -log2 ((double) (1 / (float) (1 << n)));
	 matrix_0 [index] = (float) -log2 ((double) (1 - 1 / (float) (1 << n)));
      }
}

typedef struct rle_model
{
   int16_t  count [MAXEDGES + 1];
   uint16_t total;
   uint16_t n;
   uint16_t max_domains;
   uint16_t y_index;
   int16_t  states [MAXSTATES];
   /* domain_0: qac model with max_domains == 1 (domain-pool.c:655) */
   uint16_t d0_n;
   int16_t  d0_index;
   uint16_t d0_y_index;
} rle_model_t;

/* copying only the live part of states[] is equivalent to rle_model_duplicate
   (domain-pool.c:686-705): entries >= n are never read before being rewritten */
static void
rle_copy (rle_model_t *dst, const rle_model_t *src)
{
   memcpy (dst, src, offsetof (rle_model_t, states));
   memcpy (dst->states, src->states, src->n * sizeof (int16_t));
   dst->d0_n	   = src->d0_n;
   dst->d0_index   = src->d0_index;
   dst->d0_y_index = src->d0_y_index;
}

static void
rle_init (rle_model_t *m, unsigned max_domains) /* domain-pool.c:655-672 */
{
   unsigned k;

   memset (m, 0, sizeof *m);
   for (k = m->total = 0; k < MAXEDGES + 1; k++, m->total++)
      m->count [k] = 1;
   m->max_domains = (uint16_t) max_domains;
}

static int
rle_append (rle_model_t *m, unsigned new_state) /* domain-pool.c:832-852, :467-484 */
{
   if (m->n >= m->max_domains)
      return 0;
   m->states [m->n] = (int16_t) new_state;
   m->n++;
   if (new_state == 0)
   {
      if (m->d0_n < 1)		/* qac_append into the 1-entry model */
      {
	 m->d0_index = 0;
	 m->d0_n     = 1;
      }
   }
   return 1;
}

/* rle_generate, domain-pool.c:707-735.  'domains' must hold n + 2 entries. */
static void
rle_generate (int16_t *domains, int y_state, const uint8_t *domain_type,
	      const rle_model_t *m)
{
   unsigned n;
   int	    y_is_domain = 0;

   if (y_state >= 0 && !(domain_type [y_state] & USE_DOMAIN_MASK))
      y_state = -1;
   memcpy (domains, m->states, m->n * sizeof (int16_t));
   for (n = 0; n < m->n; n++)
      if (domains [n] == y_state)
	 y_is_domain = 1;
   if (y_is_domain)
      domains [m->n] = -1;
   else
   {
      domains [m->n]	 = (int16_t) y_state;
      domains [m->n + 1] = -1;
   }
}

/* qac_bits (domain-pool.c:367-402) specialised to how rle_bits calls it for the DC
   model: domains = {0,-1}, used = {0,-1} (dc_used) or {-1} */
static float
d0_bits (int dc_used, int y_state, const rle_model_t *m)
{
   float bits = 0;

   if (m->d0_n > 0 && 0 != y_state)		/* states[0] == 0 */
      bits += matrix_0 [m->d0_index];
   if (y_state >= 0)
      bits += matrix_0 [m->d0_y_index];
   if (dc_used)
   {
      if (0 == y_state)				/* domains[0] == y_state */
      {
	 bits -= matrix_0 [m->d0_y_index];
	 bits += matrix_1 [m->d0_y_index];
      }
      else
      {
	 bits -= matrix_0 [m->d0_index];
	 bits += matrix_1 [m->d0_index];
      }
   }
   return bits;
}

static int
cmp_word (const void *a, const void *b)
{
   return (int) *(const int16_t *) a - (int) *(const int16_t *) b;
}

static float
rle_bits (const int16_t *domains, const int16_t *used_domains, int y_state,
	  const uint8_t *domain_type, const rle_model_t *m)
/* domain-pool.c:737-793, including the overwrite of 'bits' (quirk C2) */
{
   unsigned edge, n = 0, last;
   float    bits = 0;
   int16_t  sorted [MAXEDGES + 1];
   int	    into;

   if (y_state >= 0 && !(domain_type [y_state] & USE_DOMAIN_MASK))
      y_state = -1;
   if (used_domains)
   {
      int16_t domain;

      for (edge = n = 0; (domain = used_domains [edge]) != NO_EDGE; edge++)
	 if (domains [domain] != y_state)
	    sorted [n++] = used_domains [edge];
      if (n > 1)
	 qsort (sorted, n, sizeof (int16_t), cmp_word);
   }
   bits = (float) -log2 ((double) (m->count [n] / (float) m->total));
   if (used_domains && n && sorted [0] == 0)
      bits += d0_bits (1, y_state, m);
   else
      bits += d0_bits (0, y_state, m);

   last = 1;
   for (edge = 0; edge < n; edge++)
      if ((into = sorted [edge]) && (unsigned) m->n - 1 - last)
      {
	 bits += bits_bin_code ((unsigned) into - last, (unsigned) m->n - 1 - last);
	 last  = (unsigned) into + 1;
      }
   return bits;
}

static void
rle_update (const int16_t *domains, const int16_t *used_domains, int y_state,
	    const uint8_t *domain_type, rle_model_t *m)
/* domain-pool.c:795-830 with qac_update (:404-446) inlined for the DC model */
{
   int	    state_0 = 0, state_y = 0;
   unsigned edge    = 0;

   if (y_state >= 0 && !(domain_type [y_state] & USE_DOMAIN_MASK))
      y_state = -1;
   if (used_domains)
   {
      int16_t domain;

      for (edge = 0; (domain = used_domains [edge]) != NO_EDGE; edge++)
	 if (domains [domain] == 0)
	    state_0 = 1;
	 else if (domains [domain] == y_state)
	    state_y = 1;
   }
   m->count [edge]++;
   m->total++;

   /* qac_update (array0, array0 + (state_0 ? 0 : 1), ...) on domain_0 */
   {
      int y_is_domain = 0, used_y = 0;

      if (m->d0_n > 0)
      {
	 m->d0_index++;
	 if (0 == y_state)
	    y_is_domain = 1;
      }
      if (state_0)		/* one used domain: index 0, domains[0] == 0 */
      {
	 if (0 == y_state)
	 {
	    if (y_is_domain)
	       m->d0_index--;
	    m->d0_y_index >>= 1;
	    used_y = 1;
	 }
	 else
	 {
	    m->d0_index--;
	    m->d0_index >>= 1;
	 }
      }
      if (y_state >= 0 && !used_y)
	 m->d0_y_index++;
      if (m->d0_n > 0 && m->d0_index > 1020)
	 m->d0_index = 1020;
      if (m->d0_y_index > 1020)
	 m->d0_y_index = 1020;
   }
   if (state_y)
      m->y_index >>= 1;
   else
      m->y_index++;
   if (m->y_index > 1020)
      m->y_index = 1020;
}

/*****************************************************************************
			       coder state
*****************************************************************************/

typedef struct range
{
   unsigned x, y, image, address, level;
   float    weight [MAXEDGES + 1];
   int16_t  into [MAXEDGES + 1];
   int	    tree;
   float    err, tree_bits, matrix_bits, weights_bits;
   /* prediction (cwfa.h:46-75): identically 0 / "none" on the still-image path */
   float    mv_tree_bits, mv_coord_bits, nd_tree_bits, nd_weights_bits;
   int	    prediction;		/* range is coded as motion compensation + delta image */
   int	    mv_type, mv_fx, mv_fy, mv_bx, mv_by;	/* codec/wfa.h:62-71 */
} range_t;

typedef struct coder
{
   fo_params_t	opt;		/* clamped copy (coder.c:260-296) */
   float	price;
   unsigned	level;		/* image level */
   unsigned	products_level;
   unsigned	coeff_min_level, coeff_max_level;
   rpf_t	rpf, dc_rpf;
   const int16_t *planes [3];
   float       *pixels;		/* current lc_max block in bintree order */
   float       *images_of_state [MAXSTATES];
   float       *ip_images_state [MAXSTATES];
   float       *ip_states_state [MAXSTATES][MAXLEVEL];
   tree_model_t tree, p_tree;	/* bintree model, prediction-tree model (cwfa.h:96-97) */
   rle_model_t	pool, d_pool;	/* domain pool; pool of the delta approximations */
   coeff_t	coeff, d_coeff;
   int		d_pool_is_rle;	/* predicted frames: the delta pool is a real "rle" pool,
				   else the "constant" one (coder.c:720-725) */
   rle_model_t *ap;		/* the pool / coefficient model approximate_range works with */
   coeff_t     *ac;
   /* motion (cwfa.h:33-44, mwfa.c:86-126) */
   int		frame_type;	/* 0 intra, 1 predicted */
   unsigned	p_min_level, p_max_level, search_range;
   const int16_t *past;		/* regenerated reference frames (B frames: both) */
   const int16_t *future;
   float       *mc_backward_norms [MAXLEVEL];
   float	xbits [64], ybits [64];
   float       *mc_forward_norms [MAXLEVEL];
   fo_wfa_t    *wfa;
   fo_stats_t  *st;
   FILE	       *trace;
   unsigned	lc_calls, ipis_calls;
   jmp_buf	env;
   char	       *errbuf;
   size_t	errlen;
} coder_t;

static void
fail (coder_t *c, const char *fmt, ...)
{
   va_list ap;

   va_start (ap, fmt);
   if (c->errbuf && c->errlen)
      vsnprintf (c->errbuf, c->errlen, fmt, ap);
   va_end (ap);
   longjmp (c->env, 1);
}

static uint32_t
fbits (float f)
{
   uint32_t u;
   memcpy (&u, &f, 4);
   return u;
}

#define need_image(s, w) ((w)->domain_type [s] & (AUXILIARY_MASK | USE_DOMAIN_MASK))
#define usedomain(s, w)	 ((w)->domain_type [s] & USE_DOMAIN_MASK)

static void
clear_or_alloc (float **ptr, size_t size) /* control.c:259-271 */
{
   if (*ptr == NULL)
      *ptr = calloc (size ? size : 1, sizeof (float));
   else
      memset (*ptr, 0, size * sizeof (float));
}

/*****************************************************************************
			inner products  (codec/ip.c)
*****************************************************************************/

static float
standard_ip_image_state (unsigned address, unsigned level, unsigned domain,
			 coder_t *c) /* ip.c:268-295 */
{
   unsigned	i;
   float	ip = 0;
   const float *im = &c->pixels [address * size_of_level (level)];
   const float *st = c->images_of_state [domain] + address_of_level (level);

   for (i = size_of_level (level); i; i--)
      ip += *im++ * *st++;
   c->st->leaf_dots++;
   return ip;
}

static float
standard_ip_state_state (unsigned d1, unsigned d2, unsigned level,
			 const coder_t *c) /* ip.c:297-323 */
{
   unsigned	i;
   float	ip = 0;
   const float *s1 = c->images_of_state [d1] + address_of_level (level);
   const float *s2 = c->images_of_state [d2] + address_of_level (level);

   for (i = size_of_level (level); i; i--)
      ip += *s1++ * *s2++;
   return ip;
}

static float
get_ip_image_state (unsigned image, unsigned address, unsigned level,
		    unsigned domain, coder_t *c) /* ip.c:46-70 */
{
   if (level <= (unsigned) c->opt.images_level)
      return standard_ip_image_state (address, level, domain, c);
   return c->ip_images_state [domain][image];
}

static float
get_ip_state_state (unsigned d1, unsigned d2, unsigned level,
		    const coder_t *c) /* ip.c:156-182 */
{
   c->st->ipss_lookups++;
   if (level <= (unsigned) c->opt.images_level)
      return standard_ip_state_state (d1, d2, level, c);
   if (d2 < d1)
      return c->ip_states_state [d1][level][d2];
   return c->ip_states_state [d2][level][d1];
}

static void
compute_ip_images_state (unsigned image, unsigned address, unsigned level,
			 unsigned n, unsigned from, coder_t *c) /* ip.c:72-154 */
{
   const fo_wfa_t *w  = c->wfa;
   const unsigned  il = (unsigned) c->opt.images_level;
   unsigned	   state, label;

   if (level <= il)
      return;
   if (level > il + 1)
      compute_ip_images_state (MAXLABELS * image + 1, address * MAXLABELS,
			       level - 1, MAXLABELS * n, from, c);
   for (label = 0; label < MAXLABELS; label++)
      for (state = from; state < w->states; state++)
	 if (need_image (state, w))
	 {
	    unsigned edge, count;
	    int	     domain;
	    float   *dst, *src;

	    if ((domain = w->tree [state][label]) != RANGE)
	    {
	       dst = c->ip_images_state [state] + image;
	       if (level > il + 1)
	       {
		  src = c->ip_images_state [domain] + image * MAXLABELS + label + 1;
		  for (count = n; count; count--, src += MAXLABELS)
		     *dst++ += *src;
	       }
	       else
	       {
		  unsigned newadr = address * MAXLABELS + label;

		  for (count = n; count; count--, newadr += MAXLABELS)
		     *dst++ += standard_ip_image_state (newadr, level - 1,
							(unsigned) domain, c);
	       }
	    }
	    for (edge = 0; (domain = w->into [state][label][edge]) != NO_EDGE;
		 edge++)
	    {
	       float weight = w->weight [state][label][edge];

	       dst = c->ip_images_state [state] + image;
	       if (level > il + 1)
	       {
		  src = c->ip_images_state [domain] + image * MAXLABELS + label + 1;
		  for (count = n; count; count--, src += MAXLABELS)
		     *dst++ += *src * weight;
	       }
	       else
	       {
		  unsigned newadr = address * MAXLABELS + label;

		  for (count = n; count; count--, newadr += MAXLABELS)
		     *dst++ += weight * standard_ip_image_state (newadr, level - 1,
								 (unsigned) domain, c);
	       }
	    }
	 }
}

static void
compute_ip_states_state (unsigned from, unsigned to, coder_t *c) /* ip.c:184-260 */
{
   const fo_wfa_t *w = c->wfa;
   unsigned	   level, s1, s2;

   for (level = (unsigned) c->opt.images_level + 1;
	level <= (unsigned) c->opt.lc_max_level; level++)
      for (s1 = from; s1 <= to; s1++)
	 for (s2 = 0; s2 <= s1; s2++)
	    if (need_image (s2, w))
	    {
	       unsigned label;
	       float	ip = 0;

	       for (label = 0; label < MAXLABELS; label++)
	       {
		  int	   d1, d2;
		  unsigned e1, e2;
		  float	   sum, w2;

		  if ((d1 = w->tree [s1][label]) != RANGE)
		  {
		     sum = 0;
		     if ((d2 = w->tree [s2][label]) != RANGE)
			sum = get_ip_state_state ((unsigned) d1, (unsigned) d2,
						  level - 1, c);
		     for (e2 = 0; (d2 = w->into [s2][label][e2]) != NO_EDGE; e2++)
		     {
			w2   = w->weight [s2][label][e2];
			sum += w2 * get_ip_state_state ((unsigned) d1, (unsigned) d2,
							level - 1, c);
		     }
		     ip += sum;
		  }
		  for (e1 = 0; (d1 = w->into [s1][label][e1]) != NO_EDGE; e1++)
		  {
		     float w1 = w->weight [s1][label][e1];

		     sum = 0;
		     if ((d2 = w->tree [s2][label]) != RANGE)
			sum = get_ip_state_state ((unsigned) d1, (unsigned) d2,
						  level - 1, c);
		     for (e2 = 0; (d2 = w->into [s2][label][e2]) != NO_EDGE; e2++)
		     {
			w2   = w->weight [s2][label][e2];
			sum += w2 * get_ip_state_state ((unsigned) d1, (unsigned) d2,
							level - 1, c);
		     }
		     ip += w1 * sum;
		  }
	       }
	       c->ip_states_state [s1][level][s2] = ip;
	    }
}

/*****************************************************************************
		    state bookkeeping  (codec/control.c, wfalib.c)
*****************************************************************************/

static void
compute_images (unsigned from, unsigned to, coder_t *c) /* control.c:205-257 */
{
   const fo_wfa_t *w = c->wfa;
   unsigned	   label, level, state;

   for (level = 1; level <= (unsigned) c->opt.images_level; level++)
      for (state = from; state <= to; state++)
	 for (label = 0; label < MAXLABELS; label++)
	 {
	    float   *dst, *src;
	    unsigned edge, n;
	    int	     domain;

	    if ((domain = w->tree [state][label]) != RANGE)
	    {
	       dst = c->images_of_state [state] + address_of_level (level)
		     + label * size_of_level (level - 1);
	       src = c->images_of_state [domain] + address_of_level (level - 1);
	       memcpy (dst, src, size_of_level (level - 1) * sizeof (float));
	    }
	    for (edge = 0; (domain = w->into [state][label][edge]) != NO_EDGE;
		 edge++)
	    {
	       float weight = w->weight [state][label][edge];

	       dst = c->images_of_state [state] + address_of_level (level)
		     + label * size_of_level (level - 1);
	       src = c->images_of_state [domain] + address_of_level (level - 1);
	       for (n = size_of_level (level - 1); n; n--)
		  *dst++ += *src++ * weight;
	    }
	 }
}

static float
compute_final_distribution (unsigned state, const fo_wfa_t *w) /* wfalib.c:154-180 */
{
   unsigned label, edge;
   float    final = 0;
   int	    domain;

   for (label = 0; label < MAXLABELS; label++)
   {
      if ((domain = w->tree [state][label]) != RANGE)
	 final += w->final_distribution [domain];
      for (edge = 0; (domain = w->into [state][label][edge]) != NO_EDGE; edge++)
	 final += w->weight [state][label][edge] * w->final_distribution [domain];
   }
   return final / MAXLABELS;
}

static void
append_edge (unsigned from, unsigned into, float weight, unsigned label,
	     fo_wfa_t *w) /* wfalib.c:233-274 */
{
   unsigned new, edge;

   for (new = 0; (w->into [from][label][new] != NO_EDGE
		  && w->into [from][label][new] < (int) into); new++)
      ;
   for (edge = new; w->into [from][label][edge] != NO_EDGE; edge++)
      ;
   for (edge++; edge != new; edge--)
   {
      w->into [from][label][edge]   = w->into [from][label][edge - 1];
      w->weight [from][label][edge] = w->weight [from][label][edge - 1];
   }
   w->into [from][label][edge]	 = (int16_t) into;
   w->weight [from][label][edge] = weight;
}

static void
remove_states (unsigned from, fo_wfa_t *w) /* wfalib.c:276-310 */
{
   unsigned state, label;

   for (state = from; state < w->states; state++)
   {
      for (label = 0; label < MAXLABELS; label++)
      {
	 w->into [state][label][0] = NO_EDGE;
	 w->tree [state][label]	   = RANGE;
	 w->y_state [state][label] = RANGE;
	 w->mv_type [state][label] = 0;
	 w->mv_fx [state][label]   = w->mv_fy [state][label] = 0;
	 w->mv_bx [state][label]   = w->mv_by [state][label] = 0;
	 w->x [state][label]	   = w->y [state][label] = 0;	/* (not in the reference: the virtual
								   states of a colour frame never set them) */
      }
      w->domain_type [state] = 0;
      w->delta_state [state] = 0;
   }
   w->states = from;
}

static void
trace_state (coder_t *c, unsigned state, int auxiliary, unsigned level, float final)
{
   const fo_wfa_t *w = c->wfa;

   if (!c->trace)
      return;
   fprintf (c->trace, "st %u %d %u %08x %d %d\n", state, auxiliary, level,
	    fbits (final), (int) w->tree [state][0], (int) w->tree [state][1]);
   if (!auxiliary && state < 12)
   {
      unsigned i, l, t;

      fprintf (c->trace, "img %u", state);
      for (i = 0; i < size_of_tree ((unsigned) c->opt.images_level); i++)
	 fprintf (c->trace, " %08x", fbits (c->images_of_state [state][i]));
      fprintf (c->trace, "\n");
      for (l = (unsigned) c->opt.images_level + 1;
	   l <= (unsigned) c->opt.lc_max_level; l++)
      {
	 fprintf (c->trace, "ipss %u %u", state, l);
	 for (t = 0; t <= state; t++)
	    fprintf (c->trace, " %08x",
		     need_image (t, w) ? fbits (c->ip_states_state [state][l][t]) : 0);
	 fprintf (c->trace, "\n");
      }
   }
}

static void
append_state (int auxiliary, float final, unsigned level_of_state,
	      coder_t *c) /* control.c:48-131 */
{
   fo_wfa_t *w = c->wfa;
   unsigned  s = w->states, level;

   w->final_distribution [s] = final;
   w->level_of_state [s]     = (uint8_t) level_of_state;
   if (!auxiliary)
   {
      w->domain_type [s] = USE_DOMAIN_MASK;
      clear_or_alloc (&c->images_of_state [s],
		      size_of_tree ((unsigned) c->opt.images_level));
      for (level = (unsigned) c->opt.images_level + 1;
	   level <= (unsigned) c->opt.lc_max_level; level++)
	 clear_or_alloc (&c->ip_states_state [s][level], s + 1);
      clear_or_alloc (&c->ip_images_state [s], size_of_tree (c->products_level));
      c->images_of_state [s][0] = final;
      compute_images (s, s, c);
      compute_ip_states_state (s, s, c);
   }
   else
   {
      w->domain_type [s] = 0;
      free (c->images_of_state [s]);
      c->images_of_state [s] = NULL;
      for (level = 0; level < MAXLEVEL; level++)
      {
	 free (c->ip_states_state [s][level]);
	 c->ip_states_state [s][level] = NULL;
      }
      free (c->ip_images_state [s]);
      c->ip_images_state [s] = NULL;
   }
   c->st->append_states++;
   trace_state (c, s, auxiliary, level_of_state, final);
   w->states++;
   if (w->states >= MAXSTATES)
      fail (c, "Maximum number of states reached!");
}

static void
append_basis_states (coder_t *c) /* control.c:133-173 + input/basis.c:76-104,126-131 */
{
   fo_wfa_t *w = c->wfa;
   unsigned  state, level;

   /* "small.fco": state 0 = constant 128; s1, s2 as in input/basis.c:126-131 */
   w->basis_states = w->states = 3;
   w->domain_type [0]	     = USE_DOMAIN_MASK;
   w->final_distribution [0] = 128;
   append_edge (0, 0, 1.0f, 0, w);
   append_edge (0, 0, 1.0f, 1, w);
   w->final_distribution [1] = 64;
   w->final_distribution [2] = 64;
   w->domain_type [1]	     = USE_DOMAIN_MASK;
   w->domain_type [2]	     = USE_DOMAIN_MASK;
   append_edge (1, 2, 0.5f, 0, w);
   append_edge (1, 2, 0.5f, 1, w);
   append_edge (1, 0, 0.5f, 1, w);
   append_edge (2, 1, 1.0f, 0, w);
   append_edge (2, 1, 1.0f, 1, w);

   for (state = 0; state < w->basis_states; state++)
   {
      clear_or_alloc (&c->images_of_state [state],
		      size_of_tree ((unsigned) c->opt.images_level));
      for (level = (unsigned) c->opt.images_level + 1;
	   level <= (unsigned) c->opt.lc_max_level; level++)
	 clear_or_alloc (&c->ip_states_state [state][level], state + 1);
      clear_or_alloc (&c->ip_images_state [state],
		      size_of_tree (c->products_level));
      c->images_of_state [state][0] = w->final_distribution [state];
      w->level_of_state [state]	    = (uint8_t) -1;
   }
   compute_images (0, w->basis_states - 1, c);
   compute_ip_states_state (0, w->basis_states - 1, c);
}

/*****************************************************************************
		     matching pursuit  (codec/approx.c)
*****************************************************************************/

typedef struct mp
{
   int16_t exclude [MAXEDGES];
   int16_t indices [MAXEDGES + 1];
   int16_t into [MAXEDGES + 1];
   float   weight [MAXEDGES];
   float   matrix_bits, weights_bits, err, costs;
} mp_t;

/* file-static work arrays of approx.c:279-305 */
static float   norm_ortho_vector [MAXSTATES];
static float   ip_image_ortho_vector [MAXEDGES];
static float   ip_domain_ortho_vector [MAXSTATES][MAXEDGES];
static float   rem_denominator [MAXSTATES];
static float   rem_numerator [MAXSTATES];
static uint8_t used [MAXSTATES];

static void
orthogonalize (unsigned index, unsigned n, unsigned level, float min_norm,
	       const int16_t *domain_blocks, coder_t *c) /* approx.c:644-699 */
{
   unsigned domain;

   ip_image_ortho_vector [n] = rem_numerator [index];
   norm_ortho_vector [n]     = rem_denominator [index];
   c->st->ortho_steps++;

   for (domain = 0; domain_blocks [domain] >= 0; domain++)
      if (!used [domain])
      {
	 unsigned k;
	 float	  tmp = get_ip_state_state ((unsigned) domain_blocks [index],
					    (unsigned) domain_blocks [domain],
					    level, c);

	 for (k = 0; k < n; k++)
	    tmp -= ip_domain_ortho_vector [domain][k] / norm_ortho_vector [k]
		   * ip_domain_ortho_vector [index][k];
	 ip_domain_ortho_vector [domain][n] = tmp;
	 rem_denominator [domain] -= (tmp * tmp) / norm_ortho_vector [n];
	 rem_numerator [domain]	  -= ip_image_ortho_vector [n]
				     / norm_ortho_vector [n]
				     * ip_domain_ortho_vector [domain][n];
	 if (rem_denominator [domain] / size_of_level (level) < min_norm)
	    used [domain] = 1;
      }
}

static void
matching_pursuit (mp_t *mp, int full_search, float price, unsigned max_edges,
		  int y_state, const range_t *range, coder_t *c)
/* approx.c:317-642 */
{
   const fo_wfa_t *w = c->wfa;
   unsigned	   n, domain, best_n = 0;
   int		   index;
   float	   norm, additional_bits;
   const float	   min_norm = 2e-3f;
   const unsigned  size	    = size_of_level (range->level);
   static int16_t  domain_blocks [MAXSTATES + 2];

   c->st->mp_calls++;
   rle_generate (domain_blocks, y_state, w->domain_type, c->ap);
   for (domain = 0; domain_blocks [domain] >= 0; domain++)
   {
      used [domain] = 0;
      rem_denominator [domain]
	 = get_ip_state_state ((unsigned) domain_blocks [domain],
			       (unsigned) domain_blocks [domain], range->level, c);
      if (rem_denominator [domain] / size < min_norm)
	 used [domain] = 1;
      else
	 rem_numerator [domain]
	    = get_ip_image_state (range->image, range->address, range->level,
				  (unsigned) domain_blocks [domain], c);
      if (!used [domain] && fabs ((double) rem_numerator [domain]) < min_norm)
	 used [domain] = 1;
   }
   c->st->mp_domains += domain;

   for (n = 0; mp->exclude [n] != NO_EDGE; n++)
      used [mp->exclude [n]] = 1;

   for (norm = 0, n = 0; n < size; n++)
      norm += c->pixels [range->address * size + n]
	      * c->pixels [range->address * size + n];

   additional_bits = range->tree_bits + range->mv_tree_bits + range->mv_coord_bits
		     + range->nd_tree_bits + range->nd_weights_bits;

   mp->err	    = norm;
   mp->weights_bits = 0;
   mp->matrix_bits  = rle_bits (domain_blocks, NULL, y_state, w->domain_type,
				c->ap);
   mp->costs	    = (mp->matrix_bits + mp->weights_bits + additional_bits)
		      * price + mp->err;

   n = 0;
   do
   {
      float min_matrix_bits = 0, min_weights_bits = 0, min_error = 0;
      float min_weight [MAXEDGES];
      float min_costs = full_search ? MAXCOSTS : mp->costs;

      for (index = -1, domain = 0; domain_blocks [domain] >= 0; domain++)
	 if (!used [domain])
	 {
	    float matrix_bits, weights_bits;

	    c->st->pass1++;
	    {
	       int16_t	vectors [MAXEDGES + 1], states [MAXEDGES + 1];
	       float	weights [MAXEDGES + 1];
	       unsigned i, k;

	       for (i = 0, k = 0; k < n; k++)
		  if (mp->weight [k] != 0)
		  {
		     vectors [i] = mp->indices [k];
		     states [i]	 = domain_blocks [vectors [i]];
		     weights [i] = mp->weight [k];
		     i++;
		  }
	       vectors [i]     = (int16_t) domain;
	       states [i]      = domain_blocks [domain];
	       weights [i]     = 0.5f;
	       vectors [i + 1] = -1;
	       states [i + 1]  = -1;

	       weights_bits = aac_bits (weights, states, range->level, c->ac);
	       matrix_bits  = rle_bits (domain_blocks, vectors, y_state,
					w->domain_type, c->ap);
	    }
	    if (((matrix_bits + weights_bits + additional_bits) * price
		 + mp->err
		 - (rem_numerator [domain] * rem_numerator [domain])
		 / rem_denominator [domain]) < min_costs)
	    {
	       unsigned k;
	       int	l;
	       float	m_bits, w_bits, costs, m_err;
	       float	r [MAXEDGES], f [MAXEDGES];
	       int	v [MAXEDGES];

	       c->st->pass2++;
	       f [n] = rem_numerator [domain] / rem_denominator [domain];
	       v [n] = (int) domain;
	       for (k = 0; k < n; k++)
	       {
		  f [k] = ip_image_ortho_vector [k] / norm_ortho_vector [k];
		  v [k] = mp->indices [k];
	       }
	       for (l = (int) n; l >= 0; l--)
	       {
		  const rpf_t *rpf = domain_blocks [v [l]] ? &c->ac->rpf
							   : &c->ac->dc_rpf;

		  r [l] = f [l] = btor (rtob (f [l], rpf), rpf);
		  for (k = 0; k < (unsigned) l; k++)
		     f [k] -= f [l] * ip_domain_ortho_vector [v [l]][k]
			      / norm_ortho_vector [k];
	       }
	       {
		  int16_t vectors [MAXEDGES + 1], states [MAXEDGES + 1];
		  float	  weights [MAXEDGES + 1];
		  int	  i;

		  for (i = 0, k = 0; k <= n; k++)
		     if (f [k] != 0)
		     {
			vectors [i] = (int16_t) v [k];
			states [i]  = domain_blocks [v [k]];
			weights [i] = f [k];
			i++;
		     }
		  vectors [i] = -1;
		  states [i]  = -1;
		  w_bits = aac_bits (weights, states, range->level, c->ac);
		  m_bits = rle_bits (domain_blocks, vectors, y_state,
				     w->domain_type, c->ap);
	       }
	       /* the <v_l, o_n> loop of approx.c:554-569 only writes entries [..][n]
		  that are never read again; it is kept because it is
		  cheap here and keeps the work arrays bit-identical to the reference's */
	       for (l = 0; (unsigned) l <= n; l++)
	       {
		  float a = get_ip_state_state ((unsigned) domain_blocks [v [l]],
						(unsigned) domain_blocks [domain],
						range->level, c);

		  for (k = 0; k < n; k++)
		     a -= ip_domain_ortho_vector [v [l]][k] / norm_ortho_vector [k]
			  * ip_domain_ortho_vector [domain][k];
		  ip_domain_ortho_vector [v [l]][n] = a;
	       }
	       norm_ortho_vector [n]	 = rem_denominator [domain];
	       ip_image_ortho_vector [n] = rem_numerator [domain];

	       for (k = 0; k <= n; k++)
		  for (l = (int) k + 1; (unsigned) l <= n; l++)
		     r [k] += ip_domain_ortho_vector [v [l]][k] * r [l]
			      / norm_ortho_vector [k];
	       m_err = norm;
	       for (k = 0; k <= n; k++)
		  m_err += (r [k] * r [k]) * norm_ortho_vector [k]
			   - 2 * r [k] * ip_image_ortho_vector [k];

	       costs = (m_bits + w_bits + additional_bits) * price + m_err;
	       if (costs < min_costs)
	       {
		  index		   = (int) domain;
		  min_costs	   = costs;
		  min_matrix_bits  = m_bits;
		  min_weights_bits = w_bits;
		  min_error	   = m_err;
		  for (k = 0; k <= n; k++)
		     min_weight [k] = f [k];
	       }
	    }
	 }

      if (index >= 0)
      {
	 if (min_costs < mp->costs)
	 {
	    unsigned k;

	    mp->costs	     = min_costs;
	    mp->err	     = min_error;
	    mp->matrix_bits  = min_matrix_bits;
	    mp->weights_bits = min_weights_bits;
	    for (k = 0; k <= n; k++)
	       mp->weight [k] = min_weight [k];
	    best_n = n + 1;
	 }
	 mp->indices [n] = (int16_t) index;
	 mp->into [n]	 = domain_blocks [index];
	 used [index]	 = 1;
	 orthogonalize ((unsigned) index, n, range->level, min_norm,
			domain_blocks, c);
	 n++;
      }
   }
   while (n < max_edges && index >= 0);

   mp->indices [best_n] = NO_EDGE;
   mp->costs = (mp->matrix_bits + mp->weights_bits + additional_bits) * price
	       + mp->err;
}

static int
is_overflow_weight (float weight, int index, const coder_t *c) /* approx.c:172-176 */
{
   const rpf_t *rpf = index ? &c->ac->rpf : &c->ac->dc_rpf;

   return weight == btor (rtob (200, rpf), rpf)
	  || weight == btor (rtob (-200, rpf), rpf);
}

static float
approximate_range (float max_costs, float price, int max_edges, int y_state,
		   range_t *range, coder_t *c) /* approx.c:74-271 */
{
   const fo_wfa_t *w = c->wfa;
   mp_t		   mp;

   memset (&mp, 0, sizeof mp);	/* the reference leaves it uninitialised (quirk C11) */
   mp.exclude [0] = NO_EDGE;
   matching_pursuit (&mp, c->opt.full_search, price, (unsigned) max_edges,
		     y_state, range, c);

   if (c->opt.second_domain_block)
   {
      mp_t tmp = mp;

      tmp.exclude [0] = tmp.indices [0];
      tmp.exclude [1] = NO_EDGE;
      matching_pursuit (&tmp, c->opt.full_search, price, (unsigned) max_edges,
			y_state, range, c);
      if (tmp.costs < mp.costs)
	 mp = tmp;
   }
   if (c->opt.check_for_underflow)
   {
      int  iteration = -1;
      mp_t tmp	     = mp;

      do
      {
	 int i;

	 iteration++;
	 tmp.exclude [iteration] = NO_EDGE;
	 for (i = 0; tmp.indices [i] != NO_EDGE; i++)
	    if (tmp.weight [i] == 0)
	    {
	       tmp.exclude [iteration] = tmp.indices [i];
	       break;
	    }
	 if (tmp.exclude [iteration] != NO_EDGE)
	 {
	    tmp.exclude [iteration + 1] = NO_EDGE;
	    matching_pursuit (&tmp, c->opt.full_search, price,
			      (unsigned) max_edges, y_state, range, c);
	    if (tmp.costs < mp.costs)
	       mp = tmp;
	 }
      }
      while (tmp.exclude [iteration] != NO_EDGE && iteration < MAXEDGES - 1);
   }
   if (c->opt.check_for_overflow)
   {
      int  iteration = -1;
      mp_t tmp	     = mp;

      do
      {
	 int i;

	 iteration++;
	 tmp.exclude [iteration] = NO_EDGE;
	 for (i = 0; tmp.indices [i] != NO_EDGE; i++)
	    if (is_overflow_weight (tmp.weight [i], tmp.indices [i], c))
	    {
	       tmp.exclude [iteration] = tmp.indices [i];
	       break;
	    }
	 if (tmp.exclude [iteration] != NO_EDGE)
	 {
	    tmp.exclude [iteration + 1] = NO_EDGE;
	    matching_pursuit (&tmp, c->opt.full_search, price,
			      (unsigned) max_edges, y_state, range, c);
	    if (tmp.costs < mp.costs)
	       mp = tmp;
	 }
      }
      while (tmp.exclude [iteration] != NO_EDGE && iteration < MAXEDGES - 1);
   }

   if (mp.costs < max_costs)
   {
      int	     edge, new_index = 0, old_index;
      static int16_t domain_blocks [MAXSTATES + 2];

      for (old_index = 0; mp.indices [old_index] != NO_EDGE; old_index++)
	 if (mp.weight [old_index] != 0)
	 {
	    mp.indices [new_index] = mp.indices [old_index];
	    mp.into [new_index]	   = mp.into [old_index];
	    mp.weight [new_index]  = mp.weight [old_index];
	    new_index++;
	 }
      mp.indices [new_index] = NO_EDGE;
      mp.into [new_index]    = NO_EDGE;

      rle_generate (domain_blocks, y_state, w->domain_type, c->ap);
      rle_update (domain_blocks, mp.indices, y_state, w->domain_type, c->ap);
      aac_update (mp.weight, mp.into, range->level, c->ac);

      for (edge = 0; mp.indices [edge] != NO_EDGE; edge++)
      {
	 range->into [edge]   = mp.into [edge];
	 range->weight [edge] = mp.weight [edge];
      }
      range->into [edge]  = NO_EDGE;
      range->matrix_bits  = mp.matrix_bits;
      range->weights_bits = mp.weights_bits;
      range->err	  = mp.err;
      c->st->accepted++;
   }
   else
   {
      range->into [0] = NO_EDGE;
      mp.costs	      = MAXCOSTS;
   }
   return mp.costs;
}

/*****************************************************************************
		   bintree subdivision  (codec/subdivide.c)
*****************************************************************************/

static void
cut_to_bintree (float *dst, const int16_t *src, unsigned src_width,
		unsigned src_height, unsigned x0, unsigned y0, unsigned width,
		unsigned height) /* subdivide.c:504-541 */
{
   const unsigned mask01 = 0x555555, mask10 = 0xaaaaaa;
   unsigned	  x, y, xmask, ymask;

   ymask = 0;
   for (y = y0; y < y0 + height; y++, ymask = (ymask + mask10 + 1) & mask01)
   {
      xmask = 0;
      for (x = x0; x < x0 + width; x++, xmask = (xmask + mask01 + 1) & mask10)
	 if (y >= src_height || x >= src_width)
	    dst [xmask | ymask] = 0;
	 else
	    dst [xmask | ymask] = (float) (src [y * src_width + x] / 16);
   }
}

static void
init_range (range_t *range, unsigned band, coder_t *c) /* subdivide.c:612-644 */
{
   const fo_wfa_t *w = c->wfa;
   unsigned	   state, nstates = 0;

   for (state = 0; state < w->states; state++)
      if (need_image (state, w))
      {
	 memset (c->ip_images_state [state], 0,
		 size_of_tree (c->products_level) * sizeof (float));
	 nstates++;
      }
   cut_to_bintree (c->pixels, c->planes [band], (unsigned) c->opt.width,
		   (unsigned) c->opt.height, range->x, range->y,
		   width_of_level (range->level), height_of_level (range->level));
   range->address = range->image = 0;
   compute_ip_images_state (0, 0, range->level, 1, 0, c);

   c->st->blocks++;
   c->st->ip_bytes += 4ull * size_of_level (range->level)
		      + 4ull * (size_of_tree ((unsigned) c->opt.images_level)
				+ size_of_tree (c->products_level)) * nstates;
   if (c->trace && c->ipis_calls < 4)
   {
      unsigned i;

      fprintf (c->trace, "pix %u", c->ipis_calls);
      for (i = 0; i < size_of_level (range->level); i++)
	 fprintf (c->trace, " %d", (int) c->pixels [i]);
      fprintf (c->trace, "\n");
      for (state = 0; state < w->states && state < 40; state++)
	 if (need_image (state, w))
	 {
	    fprintf (c->trace, "ipis %u %u", c->ipis_calls, state);
	    for (i = 0; i < size_of_tree (c->products_level); i++)
	       fprintf (c->trace, " %08x", fbits (c->ip_images_state [state][i]));
	    fprintf (c->trace, "\n");
	 }
      c->ipis_calls++;
   }
}

static void
init_new_state (int auxiliary_state, int delta, range_t *range, const range_t *child,
		const int *y_state, coder_t *c) /* subdivide.c:549-610 */
{
   fo_wfa_t *w = c->wfa;
   unsigned  label, edge;
   int	     state_is_domain = 0;

   /* options.delta_domains = options.normal_domains = YES (options.c:91-92): every state
      that may be a domain enters both pools.  The delta pool of an intra frame is the
      "constant" pool whose append() always says YES (domain-pool.c:962-967,
      coder.c:720-725) */
   if (!auxiliary_state)
   {
      state_is_domain = rle_append (&c->pool, w->states);
      state_is_domain = (c->d_pool_is_rle ? rle_append (&c->d_pool, w->states) : 1)
			|| state_is_domain;
   }

   range->into [0] = NO_EDGE;
   range->tree	   = (int) w->states;
   for (label = 0; label < MAXLABELS; label++)
   {
      w->tree [w->states][label]    = (int16_t) child [label].tree;
      w->y_state [w->states][label] = (int16_t) y_state [label];
      w->mv_type [w->states][label] = (int8_t) child [label].mv_type;
      w->mv_fx [w->states][label]   = (int8_t) child [label].mv_fx;
      w->mv_fy [w->states][label]   = (int8_t) child [label].mv_fy;
      w->mv_bx [w->states][label]   = (int8_t) child [label].mv_bx;
      w->mv_by [w->states][label]   = (int8_t) child [label].mv_by;
      w->x [w->states][label]	    = (uint16_t) child [label].x;
      w->y [w->states][label]	    = (uint16_t) child [label].y;
      /* append_transitions, control.c:175-197 */
      w->y_column [w->states][label] = 0;
      for (edge = 0; child [label].into [edge] != NO_EDGE; edge++)
      {
	 append_edge (w->states, (unsigned) child [label].into [edge],
		      child [label].weight [edge], label, w);
	 if (child [label].into [edge] == w->y_state [w->states][label])
	    w->y_column [w->states][label] = 1;
      }
   }
   w->delta_state [w->states] = (uint8_t) delta;
   append_state (!state_is_domain, compute_final_distribution (w->states, w),
		 range->level, c);
}

static float subdivide (float max_costs, unsigned band, int y_state, range_t *range,
			coder_t *c, int prediction, int delta);

/*****************************************************************************
		 motion search  (codec/mwfa.c, codec/prediction.c)
*****************************************************************************/

static void extract_mc_block (int16_t *mcblock, unsigned width, unsigned height,
			      const int16_t *reference, unsigned ref_width, int half_pixel,
			      unsigned xo, unsigned yo, int mx, int my);

/* MPEG's code lengths of the vector components (mwfa.c:40-52, second column) */
static const unsigned char mv_code_length [33] =
{11, 11, 11, 11, 11, 11, 10, 10, 10, 8, 8, 8, 7, 5, 4, 3, 1, 3, 4, 5, 7, 8, 8, 8, 10, 10, 10,
 11, 11, 11, 11, 11, 11};

static unsigned
norms_size (const coder_t *c)	/* full-pixel search: (2 * search_range)^2 vectors */
{
   return 4 * c->search_range * c->search_range;
}

/* clear_norms_table (prediction.c:195-211) */
static void
clear_norms_table (unsigned level, coder_t *c)
{
   if (level > c->p_min_level)
   {
      memset (c->mc_forward_norms [level], 0, norms_size (c) * sizeof (float));
      memset (c->mc_backward_norms [level], 0, norms_size (c) * sizeof (float));
   }
}

/* update_norms_table (prediction.c:213-238): a level's norms are the sums of its children's */
static void
update_norms_table (unsigned level, coder_t *c)
{
   if (level > c->p_min_level)
   {
      for (unsigned index = 0; index < norms_size (c); index++)
	 c->mc_forward_norms [level][index] += c->mc_forward_norms [level - 1][index];
      if (c->frame_type == 2)
	 for (unsigned index = 0; index < norms_size (c); index++)
	    c->mc_backward_norms [level][index] += c->mc_backward_norms [level - 1][index];
   }
}

/* mcpe_norm (mwfa.c:651-684) over get_mcpe (:604-649), forward prediction: the squared
   norm of (original - displaced reference) / 16, summed in row order in fp32 */
static float
mcpe_norm (const coder_t *c, unsigned x0, unsigned y0, unsigned width, unsigned height,
	   const int16_t *mcblock, const int16_t *mcblock2)
{
   const int16_t *o    = c->planes [0] + (size_t) y0 * (unsigned) c->opt.width + x0;
   float	  norm = 0;

   for (unsigned y = 0; y < height; y++)
      for (unsigned x = 0; x < width; x++)
      {
	 /* get_mcpe (mwfa.c:604-649): one reference block, or the mean of two (truncating) */
	 const int     ref = mcblock2 ? (mcblock [y * width + x] + mcblock2 [y * width + x]) / 2
				      : mcblock [y * width + x];
	 const int16_t d   = (int16_t) (o [(size_t) y * (unsigned) c->opt.width + x] - ref);
	 const int     q   = d / 16;

	 norm += (float) (q * q);
      }
   return norm;
}

/* fill_norms_table (mwfa.c:544-602), P frames */
static void
fill_norms_table (unsigned x0, unsigned y0, unsigned level, coder_t *c)
{
   const int	  sr	 = (int) c->search_range;
   const unsigned width	 = width_of_level (level), height = height_of_level (level);
   int16_t	 *mcblock = malloc ((size_t) width * height * sizeof (int16_t));
   unsigned	  index	 = 0;

   for (int my = -sr; my < sr; my++)
      for (int mx = -sr; mx < sr; mx++, index++)
      {
	 if ((int) x0 + mx < 0 || x0 + mx + width > (unsigned) c->opt.width
	     || (int) y0 + my < 0 || y0 + my + height > (unsigned) c->opt.height)
	 {
	    c->mc_forward_norms [level][index]	= 0.0f;
	    c->mc_backward_norms [level][index] = 0.0f;
	 }
	 else
	 {
	    extract_mc_block (mcblock, width, height, c->past, (unsigned) c->opt.width, 0,
			      x0, y0, mx, my);
	    c->mc_forward_norms [level][index] = mcpe_norm (c, x0, y0, width, height, mcblock, NULL);
	    if (c->frame_type == 2)
	    {
	       extract_mc_block (mcblock, width, height, c->future, (unsigned) c->opt.width, 0,
				 x0, y0, mx, my);
	       c->mc_backward_norms [level][index] = mcpe_norm (c, x0, y0, width, height, mcblock, NULL);
	    }
	 }
      }
   free (mcblock);
}

/* find_best_mv (mwfa.c:686-795), full-pixel search */
static float
find_best_mv (float price, unsigned x0, unsigned y0, unsigned width, unsigned height,
	      float *bits, int *mx, int *my, const float *mc_norms, const coder_t *c)
{
   const int sr	      = (int) c->search_range;
   float     mincosts = MAXCOSTS;
   unsigned  index    = 0;

   *mx = *my = 0;
   for (int y = -sr; y < sr; y++)
      for (int x = -sr; x < sr; x++, index++)
	 if ((int) x0 + x >= 0 && (int) y0 + y >= 0
	     && x0 + x + width <= (unsigned) c->opt.width
	     && y0 + y + height <= (unsigned) c->opt.height)
	 {
	    const float costs = mc_norms [index] + (c->xbits [x + sr] + c->ybits [y + sr]) * price;

	    if (costs < mincosts)
	    {
	       mincosts = costs;
	       *mx	= x;
	       *my	= y;
	    }
	 }
   *bits = c->xbits [*mx + sr] + c->ybits [*my + sr];
   return mincosts;
}

/* saved state data (prediction.c:47-70) */
typedef struct state_data
{
   float    final_distribution;
   uint8_t  level_of_state, domain_type;
   float   *images_of_state, *inner_products, *ip_states_state [MAXLEVEL];
   int16_t  tree [MAXLABELS], y_state [MAXLABELS], into [MAXLABELS][MAXEDGES + 1];
   uint8_t  y_column [MAXLABELS];
   int8_t   mv_type [MAXLABELS], mv_fx [MAXLABELS], mv_fy [MAXLABELS], mv_bx [MAXLABELS], mv_by [MAXLABELS];
   uint16_t x [MAXLABELS], y [MAXLABELS];
   float    weight [MAXLABELS][MAXEDGES + 1];
} state_data_t;

/* store_state_data (prediction.c:502-560): move the states from..to aside.  The delta
   flag of a state is NOT part of the saved data (nor of the restored data below) */
static state_data_t *
store_state_data (unsigned from, unsigned to, coder_t *c)
{
   fo_wfa_t	*w = c->wfa;
   state_data_t *data;

   if (to + 1 <= from)
      return NULL;
   data = calloc (to - from + 1, sizeof *data);
   for (unsigned state = from; state <= to; state++)
   {
      state_data_t *sd = &data [state - from];

      sd->final_distribution = w->final_distribution [state];
      sd->level_of_state     = w->level_of_state [state];
      sd->domain_type	     = w->domain_type [state];
      sd->images_of_state    = c->images_of_state [state];
      sd->inner_products     = c->ip_images_state [state];
      w->domain_type [state]	 = 0;
      c->images_of_state [state] = NULL;
      c->ip_images_state [state] = NULL;
      for (unsigned label = 0; label < MAXLABELS; label++)
      {
	 sd->tree [label]     = w->tree [state][label];
	 sd->y_state [label]  = w->y_state [state][label];
	 sd->y_column [label] = w->y_column [state][label];
	 sd->mv_type [label]  = w->mv_type [state][label];
	 sd->mv_fx [label]    = w->mv_fx [state][label];
	 sd->mv_fy [label]    = w->mv_fy [state][label];
	 sd->mv_bx [label]    = w->mv_bx [state][label];
	 sd->mv_by [label]    = w->mv_by [state][label];
	 sd->x [label]	      = w->x [state][label];
	 sd->y [label]	      = w->y [state][label];
	 memcpy (sd->weight [label], w->weight [state][label], sizeof sd->weight [label]);
	 memcpy (sd->into [label], w->into [state][label], sizeof sd->into [label]);
	 w->into [state][label][0] = NO_EDGE;
	 w->tree [state][label]	   = RANGE;
	 w->y_state [state][label] = RANGE;
      }
      for (unsigned level = (unsigned) c->opt.images_level + 1;
	   level <= (unsigned) c->opt.lc_max_level; level++)
      {
	 sd->ip_states_state [level]	   = c->ip_states_state [state][level];
	 c->ip_states_state [state][level] = NULL;
      }
   }
   return data;
}

/* restore_state_data (prediction.c:562-625) */
static void
restore_state_data (unsigned from, unsigned to, state_data_t *data, coder_t *c)
{
   fo_wfa_t *w = c->wfa;

   if (to + 1 <= from)
      return;
   for (unsigned state = from; state <= to; state++)
   {
      state_data_t *sd = &data [state - from];

      w->final_distribution [state] = sd->final_distribution;
      w->level_of_state [state]	    = sd->level_of_state;
      w->domain_type [state]	    = sd->domain_type;
      free (c->images_of_state [state]);
      c->images_of_state [state] = sd->images_of_state;
      free (c->ip_images_state [state]);
      c->ip_images_state [state] = sd->inner_products;
      for (unsigned label = 0; label < MAXLABELS; label++)
      {
	 w->tree [state][label]	    = sd->tree [label];
	 w->y_state [state][label]  = sd->y_state [label];
	 w->y_column [state][label] = sd->y_column [label];
	 w->mv_type [state][label]  = sd->mv_type [label];
	 w->mv_fx [state][label]    = sd->mv_fx [label];
	 w->mv_fy [state][label]    = sd->mv_fy [label];
	 w->mv_bx [state][label]    = sd->mv_bx [label];
	 w->mv_by [state][label]    = sd->mv_by [label];
	 w->x [state][label]	    = sd->x [label];
	 w->y [state][label]	    = sd->y [label];
	 memcpy (w->weight [state][label], sd->weight [label], sizeof sd->weight [label]);
	 memcpy (w->into [state][label], sd->into [label], sizeof sd->into [label]);
      }
      for (unsigned level = (unsigned) c->opt.images_level + 1;
	   level <= (unsigned) c->opt.lc_max_level; level++)
      {
	 free (c->ip_states_state [state][level]);
	 c->ip_states_state [state][level] = sd->ip_states_state [level];
      }
   }
   free (data);
   w->states = to + 1;
}

static void
free_state_data (unsigned from, unsigned to, state_data_t *data, coder_t *c)
{
   if (to + 1 <= from)
      return;
   for (unsigned state = from; state <= to; state++)
   {
      state_data_t *sd = &data [state - from];

      for (unsigned level = (unsigned) c->opt.images_level + 1;
	   level <= (unsigned) c->opt.lc_max_level; level++)
	 free (sd->ip_states_state [level]);
      free (sd->images_of_state);
      free (sd->inner_products);
   }
   free (data);
}

/* mc_prediction (prediction.c:262-370), P frames */
static float
mc_prediction (float max_costs, float price, unsigned band, int y_state, range_t *range,
	       coder_t *c)
{
   fo_wfa_t	 *w	 = c->wfa;
   range_t	  prange = *range;
   const unsigned width	 = width_of_level (range->level);
   const unsigned height = height_of_level (range->level);
   int16_t	 *mcpe	 = calloc ((size_t) width * height, sizeof (int16_t));
   float	  costs;

   if (prange.level == c->p_min_level)
      fill_norms_table (prange.x, prange.y, prange.level, c);
   if (c->frame_type == 2)
   {
      /* find_B_frame_mc (mwfa.c:341-542) without cross-B search (coder.c:359 sets that flag
	 from the half-pixel option): best forward vector, best backward vector, both together */
      int16_t *mcblock1 = calloc ((size_t) width * height, sizeof (int16_t));
      int16_t *mcblock2 = calloc ((size_t) width * height, sizeof (int16_t));
      float    forward_bits, backward_bits, interp_bits, forward_costs, backward_costs, interp_costs;
      int      fx, fy, bx, by, mctype;

      forward_costs = find_best_mv (price, prange.x, prange.y, width, height, &forward_bits, &fx, &fy,
				    c->mc_forward_norms [prange.level], c) + 3 * price;
      backward_costs = find_best_mv (price, prange.x, prange.y, width, height, &backward_bits, &bx, &by,
				     c->mc_backward_norms [prange.level], c) + 3 * price;
      interp_bits = forward_bits + backward_bits;
      extract_mc_block (mcblock1, width, height, c->past, (unsigned) c->opt.width, 0,
			prange.x, prange.y, fx, fy);
      extract_mc_block (mcblock2, width, height, c->future, (unsigned) c->opt.width, 0,
			prange.x, prange.y, bx, by);
      interp_costs = mcpe_norm (c, prange.x, prange.y, width, height, mcblock1, mcblock2)
		     + (interp_bits + 2) * price;
      if (forward_costs <= interp_costs)
	 mctype = forward_costs <= backward_costs ? 1 : 2;
      else
	 mctype = backward_costs <= interp_costs ? 2 : 3;
      prange.mv_type = mctype;
      if (mctype == 1)
      {
	 prange.mv_tree_bits  = 3;
	 prange.mv_coord_bits = forward_bits;
	 prange.mv_fx	      = fx;
	 prange.mv_fy	      = fy;
      }
      else if (mctype == 2)
      {
	 prange.mv_tree_bits  = 3;
	 prange.mv_coord_bits = backward_bits;
	 prange.mv_bx	      = bx;
	 prange.mv_by	      = by;
	 memcpy (mcblock1, mcblock2, (size_t) width * height * sizeof (int16_t));
      }
      else
      {
	 prange.mv_tree_bits  = 2;
	 prange.mv_coord_bits = interp_bits;
	 prange.mv_fx	      = fx;
	 prange.mv_fy	      = fy;
	 prange.mv_bx	      = bx;
	 prange.mv_by	      = by;
      }
      for (unsigned y = 0; y < height; y++)
	 for (unsigned x = 0; x < width; x++)
	 {
	    const int ref = mctype == 3 ? (mcblock1 [y * width + x] + mcblock2 [y * width + x]) / 2
					: mcblock1 [y * width + x];

	    mcpe [y * width + x]
	       = (int16_t) (c->planes [0][(size_t) (prange.y + y) * (unsigned) c->opt.width + prange.x + x] - ref);
	 }
      free (mcblock1);
      free (mcblock2);
   }
   else
   /* find_P_frame_mc (mwfa.c:301-339) */
   {
      int16_t *mcblock = calloc ((size_t) width * height, sizeof (int16_t));

      prange.mv_tree_bits = 1;
      prange.mv_type	  = 1;	/* FORWARD */
      find_best_mv (price, prange.x, prange.y, width, height, &prange.mv_coord_bits,
		    &prange.mv_fx, &prange.mv_fy, c->mc_forward_norms [prange.level], c);
      if (c->trace)
	 fprintf (c->trace, "mv %u %u %u %d %d %08x\n", prange.level, prange.x, prange.y,
		  prange.mv_fx, prange.mv_fy, fbits (prange.mv_coord_bits));
      extract_mc_block (mcblock, width, height, c->past, (unsigned) c->opt.width, 0,
			prange.x, prange.y, prange.mv_fx, prange.mv_fy);
      for (unsigned y = 0; y < height; y++)	/* get_mcpe (mwfa.c:604-649) */
	 for (unsigned x = 0; x < width; x++)
	    mcpe [y * width + x]
	       = (int16_t) (c->planes [0][(size_t) (prange.y + y) * (unsigned) c->opt.width + prange.x + x]
			    - mcblock [y * width + x]);
      free (mcblock);
   }
   costs = (prange.mv_tree_bits + prange.mv_coord_bits) * price;

   if (costs < max_costs)
   {
      const unsigned last_state = w->states - 1;
      float	    *rec_pixels = c->pixels;
      float	   **ipi	= calloc (MAXSTATES, sizeof (float *));
      float	     mvt, mvc;

      c->pixels = calloc ((size_t) width * height, sizeof (float));
      cut_to_bintree (c->pixels, mcpe, width, height, 0, 0, width, height);
      for (unsigned state = 0; state <= last_state; state++)
	 if (need_image (state, w))
	 {
	    ipi [state]		       = c->ip_images_state [state];
	    c->ip_images_state [state] = calloc (size_of_tree (c->products_level), sizeof (float));
	 }
      mvc = prange.mv_coord_bits;
      mvt = prange.mv_tree_bits;
      prange.image	     = 0;
      prange.address	     = 0;
      prange.tree_bits	     = 0;
      prange.matrix_bits     = 0;
      prange.weights_bits    = 0;
      prange.mv_coord_bits   = 0;
      prange.mv_tree_bits    = 0;
      prange.nd_weights_bits = 0;
      prange.nd_tree_bits    = 0;

      compute_ip_images_state (prange.image, prange.address, prange.level, 1, 0, c);
      costs += subdivide (max_costs - costs, band, y_state, &prange, c, 0, 1);

      if (costs < max_costs)
      {
	 const unsigned img = range->image, adr = range->address;

	 *range		      = prange;
	 range->image	      = img;
	 range->address	      = adr;
	 range->mv_coord_bits = mvc;
	 range->mv_tree_bits  = mvt;
	 range->prediction    = 1;
	 for (unsigned state = last_state + 1; state < w->states; state++)
	    if (need_image (state, w))
	       memset (c->ip_images_state [state], 0,
		       size_of_tree (c->products_level) * sizeof (float));
	 costs = (range->tree_bits + range->matrix_bits + range->weights_bits
		  + range->mv_tree_bits + range->mv_coord_bits + range->nd_tree_bits
		  + range->nd_weights_bits) * price + range->err;
      }
      else
	 costs = MAXCOSTS;
      for (unsigned state = 0; state <= last_state; state++)
	 if (need_image (state, w))
	 {
	    free (c->ip_images_state [state]);
	    c->ip_images_state [state] = ipi [state];
	 }
      free (ipi);
      free (c->pixels);
      c->pixels = rec_pixels;
   }
   else
      costs = MAXCOSTS;
   free (mcpe);
   return costs;
}

/* nd_prediction (prediction.c:371-500): an intra frame coded with `--prediction': the range is
   predicted by its DC component (state 0, weight quantised with the DC format) and the
   difference is subdivided with the delta models; taken only if the difference is split */
static int nd_prediction_on = 0;
static unsigned nd_zero_weights = 0;	/* DC weights that rounded to zero since the last fo_set_nd_prediction() */

unsigned
fo_nd_zero_weights (void)
{
   return nd_zero_weights;
}

void
fo_set_nd_prediction (int on)
{
   nd_prediction_on = on;
   if (on)
      nd_zero_weights = 0;
}

static float
nd_prediction (float max_costs, float price, unsigned band, int y_state, range_t *range,
	       coder_t *c)
{
   fo_wfa_t *w	    = c->wfa;
   range_t   lrange = *range;
   float     costs;

   {
      const float   x	  = get_ip_image_state (range->image, range->address, range->level, 0, c);
      const float   y	  = get_ip_state_state (0, 0, range->level, c);
      const float   wt	  = btor (rtob (x / y, &c->dc_rpf), &c->dc_rpf);
      const int16_t s [2] = {0, -1};

      lrange.into [0]	     = 0;
      lrange.into [1]	     = NO_EDGE;
      lrange.weight [0]	     = wt;
      lrange.mv_coord_bits   = 0;
      lrange.mv_tree_bits    = 0;
      lrange.nd_tree_bits    = tree_bits (0, lrange.level, &c->p_tree);
      lrange.nd_weights_bits = 0;
      lrange.tree_bits	     = 0;
      lrange.matrix_bits     = 0;
      /* a weight that rounds to zero has the code RPF_ZERO = -1: the reference then reads the 16-bit
	 word in front of its table of counts (coeff.c:237) -- the upper end of the allocator's chunk
	 size, zero -- and the bits are -log2 (0) = infinite: such a prediction is never taken */
      lrange.weights_bits    = rtob (wt, &c->dc_rpf) < 0 ? INFINITY
						      : aac_bits (&wt, s, range->level, &c->coeff);
      if (rtob (wt, &c->dc_rpf) < 0)
	 nd_zero_weights++;
   }
   costs = price * (lrange.weights_bits + lrange.nd_tree_bits);

   if (costs < max_costs)
   {
      const unsigned width = width_of_level (range->level), height = height_of_level (range->level);
      const unsigned last_state = w->states - 1;
      float	    *rec_pixels = c->pixels;
      float	   **ipi	= calloc (MAXSTATES, sizeof (float *));
      float	    *pixels	= calloc ((size_t) width * height, sizeof (float));
      range_t	     rrange;

      {
	 const float  dc  = -lrange.weight [0] * c->images_of_state [0][0];
	 const float *src = c->pixels + (size_t) range->address * size_of_level (range->level);

	 for (unsigned n = 0; n < width * height; n++)
	    pixels [n] = src [n] + dc;
      }
      c->pixels		     = pixels;
      rrange		     = *range;
      rrange.tree_bits	     = 0;
      rrange.matrix_bits     = 0;
      rrange.weights_bits    = 0;
      rrange.mv_coord_bits   = 0;
      rrange.mv_tree_bits    = 0;
      rrange.nd_tree_bits    = 0;
      rrange.nd_weights_bits = 0;
      rrange.image	     = 0;
      rrange.address	     = 0;
      for (unsigned state = 0; state <= last_state; state++)
	 if (need_image (state, w))
	 {
	    ipi [state]		       = c->ip_images_state [state];
	    c->ip_images_state [state] = calloc (size_of_tree (c->products_level), sizeof (float));
	 }
      compute_ip_images_state (rrange.image, rrange.address, rrange.level, 1, 0, c);
      costs += subdivide (max_costs - costs, band, y_state, &rrange, c, 0, 1);

      if (costs < max_costs && rrange.tree != RANGE)
      {
	 const unsigned img = range->image, adr = range->address;
	 unsigned	edge;

	 *range			 = rrange;
	 range->image		 = img;
	 range->address		 = adr;
	 range->nd_tree_bits	+= lrange.nd_tree_bits;
	 range->nd_weights_bits += lrange.weights_bits;
	 for (edge = 0; lrange.into [edge] != NO_EDGE; edge++)
	 {
	    range->into [edge]	 = lrange.into [edge];
	    range->weight [edge] = lrange.weight [edge];
	 }
	 range->into [edge] = NO_EDGE;
	 range->prediction  = (int) edge;
	 for (unsigned state = last_state + 1; state < w->states; state++)
	    if (need_image (state, w))
	       memset (c->ip_images_state [state], 0,
		       size_of_tree (c->products_level) * sizeof (float));
      }
      else
	 costs = MAXCOSTS;
      for (unsigned state = 0; state <= last_state; state++)
	 if (need_image (state, w))
	 {
	    free (c->ip_images_state [state]);
	    c->ip_images_state [state] = ipi [state];
	 }
      free (ipi);
      free (pixels);
      c->pixels = rec_pixels;
   }
   else
      costs = MAXCOSTS;
   return costs;
}

/*
 *  Design check for the device: with holes_mode set, predict_range keeps
 *  the split alternative's states where they are instead of moving them aside -- the
 *  prediction alternative builds its states behind them; a lost prediction drops its own
 *  states, a won one leaves the split's states as dead holes that fo_close_holes() removes
 *  at the end by a monotone renumbering.  The result must equal the reference's.
 */
static int holes_mode = 0;

void
fo_set_holes_mode (int on)
{
   holes_mode = on;
}

static float
predict_range_holes (float max_costs, float price, range_t *range, coder_t *c, unsigned band,
		     int y_state, unsigned states, const tree_model_t *tree_model,
		     const tree_model_t *p_tree_model, const rle_model_t *domain_model,
		     const rle_model_t *d_domain_model, const aac_model_t *coeff_model,
		     const aac_model_t *d_coeff_model)
{
   fo_wfa_t	*w = c->wfa;
   rle_model_t	*rec_domain_model   = malloc (sizeof (rle_model_t));
   rle_model_t	*rec_d_domain_model = malloc (sizeof (rle_model_t));
   aac_model_t	 rec_coeff_model    = c->coeff.model;
   aac_model_t	 rec_d_coeff_model  = c->d_coeff.model;
   tree_model_t	 rec_tree_model	    = c->tree;
   tree_model_t	 rec_p_tree_model   = c->p_tree;
   const unsigned rec_states	    = w->states;
   uint8_t	*rec_domain_type    = malloc (rec_states - states + 1);
   float	 costs;

   rle_copy (rec_domain_model, &c->pool);
   rle_copy (rec_d_domain_model, &c->d_pool);
   /* hide the split's states: nobody may use them as domains or needs their images */
   for (unsigned s2 = states; s2 < rec_states; s2++)
   {
      rec_domain_type [s2 - states] = w->domain_type [s2];
      w->domain_type [s2]	    = 0;
   }
   c->tree	    = *tree_model;
   c->p_tree	    = *p_tree_model;
   rle_copy (&c->pool, domain_model);
   rle_copy (&c->d_pool, d_domain_model);
   c->coeff.model   = *coeff_model;
   c->d_coeff.model = *d_coeff_model;

   costs = c->frame_type == 0 ? nd_prediction (max_costs, price, band, y_state, range, c)
			      : mc_prediction (max_costs, price, band, y_state, range, c);

   if (costs < MAXCOSTS)
   {
      /* the split's states stay behind as holes: inert, never referenced */
      for (unsigned s2 = states; s2 < rec_states; s2++)
	 for (unsigned label = 0; label < MAXLABELS; label++)
	 {
	    w->into [s2][label][0] = NO_EDGE;
	    w->tree [s2][label]	   = RANGE;
	    w->mv_type [s2][label] = 0;
	    w->level_of_state [s2] = 255;	/* marks a hole for fo_close_holes */
	 }
      costs = (range->tree_bits + range->matrix_bits + range->weights_bits
	       + range->mv_tree_bits + range->mv_coord_bits + range->nd_tree_bits
	       + range->nd_weights_bits) * price + range->err;
   }
   else
   {
      rle_copy (&c->pool, rec_domain_model);
      rle_copy (&c->d_pool, rec_d_domain_model);
      c->coeff.model   = rec_coeff_model;
      c->d_coeff.model = rec_d_coeff_model;
      c->tree	       = rec_tree_model;
      c->p_tree	       = rec_p_tree_model;
      range->prediction = 0;
      if (w->states != rec_states)
	 remove_states (rec_states, w);
      for (unsigned s2 = states; s2 < rec_states; s2++)
	 w->domain_type [s2] = rec_domain_type [s2 - states];
      costs = MAXCOSTS;
   }
   if (c->trace)
      fprintf (c->trace, "pr %u %u %u %u %u %08x %08x %u\n", range->level, range->x, range->y, states,
	       rec_states, fbits (max_costs), fbits (costs), w->states);
   free (rec_domain_model);
   free (rec_d_domain_model);
   free (rec_domain_type);
   return costs;
}

/* renumber the live states of a holes-mode automaton consecutively (what the host would do) */
void
fo_close_holes (fo_wfa_t *w)
{
   static int16_t map [MAXSTATES];
   unsigned	  n = 0;

   for (unsigned s2 = 0; s2 < w->states; s2++)
      map [s2] = (s2 >= w->basis_states && w->level_of_state [s2] == 255) ? (int16_t) -1 : (int16_t) n++;
   for (unsigned s2 = 0; s2 < w->states; s2++)
   {
      const int t = map [s2];

      if (t < 0 || (unsigned) t == s2)
	 continue;
      w->final_distribution [t] = w->final_distribution [s2];
      w->level_of_state [t]	= w->level_of_state [s2];
      w->domain_type [t]	= w->domain_type [s2];
      w->delta_state [t]	= w->delta_state [s2];
      for (unsigned label = 0; label < MAXLABELS; label++)
      {
	 w->tree [t][label]	= w->tree [s2][label];
	 w->x [t][label]	= w->x [s2][label];
	 w->y [t][label]	= w->y [s2][label];
	 w->y_state [t][label]	= w->y_state [s2][label];
	 w->y_column [t][label] = w->y_column [s2][label];
	 w->mv_type [t][label]	= w->mv_type [s2][label];
	 w->mv_fx [t][label]	= w->mv_fx [s2][label];
	 w->mv_fy [t][label]	= w->mv_fy [s2][label];
	 w->mv_bx [t][label]	= w->mv_bx [s2][label];
	 w->mv_by [t][label]	= w->mv_by [s2][label];
	 memcpy (w->into [t][label], w->into [s2][label], sizeof w->into [t][label]);
	 memcpy (w->weight [t][label], w->weight [s2][label], sizeof w->weight [t][label]);
      }
   }
   for (unsigned s2 = 0; s2 < n; s2++)
      for (unsigned label = 0; label < MAXLABELS; label++)
      {
	 if (w->tree [s2][label] != RANGE)
	    w->tree [s2][label] = map [w->tree [s2][label]];
	 for (unsigned e = 0; w->into [s2][label][e] != NO_EDGE; e++)
	    w->into [s2][label][e] = map [w->into [s2][label][e]];
      }
   w->root_state = (unsigned) map [w->root_state];
   w->states	 = n;
}

/* predict_range (prediction.c:96-191), P frames */
static float
predict_range (float max_costs, float price, range_t *range, coder_t *c, unsigned band,
	       int y_state, unsigned states, const tree_model_t *tree_model,
	       const tree_model_t *p_tree_model, const rle_model_t *domain_model,
	       const rle_model_t *d_domain_model, const aac_model_t *coeff_model,
	       const aac_model_t *d_coeff_model)
{
   fo_wfa_t	*w = c->wfa;
   rle_model_t	*rec_domain_model   = malloc (sizeof (rle_model_t));
   rle_model_t	*rec_d_domain_model = malloc (sizeof (rle_model_t));
   aac_model_t	 rec_coeff_model    = c->coeff.model;
   aac_model_t	 rec_d_coeff_model  = c->d_coeff.model;
   tree_model_t	 rec_tree_model	    = c->tree;
   tree_model_t	 rec_p_tree_model   = c->p_tree;
   unsigned	 rec_states	    = w->states;
   state_data_t *rec_state_data;
   float	 costs;

   if (holes_mode)
   {
      free (rec_domain_model);
      free (rec_d_domain_model);
      return predict_range_holes (max_costs, price, range, c, band, y_state, states, tree_model,
				  p_tree_model, domain_model, d_domain_model, coeff_model,
				  d_coeff_model);
   }
   rle_copy (rec_domain_model, &c->pool);
   rle_copy (rec_d_domain_model, &c->d_pool);
   rec_state_data = store_state_data (states, rec_states - 1, c);

   w->states	    = states;
   c->tree	    = *tree_model;
   c->p_tree	    = *p_tree_model;
   rle_copy (&c->pool, domain_model);
   rle_copy (&c->d_pool, d_domain_model);
   c->coeff.model   = *coeff_model;
   c->d_coeff.model = *d_coeff_model;

   costs = c->frame_type == 0 ? nd_prediction (max_costs, price, band, y_state, range, c)
			      : mc_prediction (max_costs, price, band, y_state, range, c);

   if (costs < MAXCOSTS)
   {
      free_state_data (states, rec_states - 1, rec_state_data, c);
      costs = (range->tree_bits + range->matrix_bits + range->weights_bits
	       + range->mv_tree_bits + range->mv_coord_bits + range->nd_tree_bits
	       + range->nd_weights_bits) * price + range->err;
   }
   else
   {
      rle_copy (&c->pool, rec_domain_model);
      rle_copy (&c->d_pool, rec_d_domain_model);
      c->coeff.model   = rec_coeff_model;
      c->d_coeff.model = rec_d_coeff_model;
      c->tree	       = rec_tree_model;
      c->p_tree	       = rec_p_tree_model;
      range->prediction = 0;
      if (w->states != states)
	 remove_states (states, w);
      restore_state_data (states, rec_states - 1, rec_state_data, c);
      costs = MAXCOSTS;
   }
   if (c->trace)
      fprintf (c->trace, "pr %u %u %u %u %u %08x %08x %u\n", range->level, range->x, range->y, states,
	       rec_states, fbits (max_costs), fbits (costs), w->states);
   free (rec_domain_model);
   free (rec_d_domain_model);
   return costs;
}

static float
subdivide (float max_costs, unsigned band, int y_state, range_t *range,
	   coder_t *c, int prediction, int delta) /* subdivide.c:60-502 */
{
   fo_wfa_t    *w = c->wfa;
   float	subdivide_costs, lincomb_costs, price;
   int		new_y_state [MAXLABELS];
   unsigned	states;
   int		try_mc, try_nd;
   rle_model_t *domain_model, *lc_domain_model, *d_domain_model, *lc_d_domain_model;
   aac_model_t	coeff_model, lc_coeff_model, d_coeff_model, lc_d_coeff_model;
   tree_model_t tree_model, p_tree_model;
   range_t	lrange, rrange, child [MAXLABELS];

   c->st->subdivide_calls++;
   range->into [0] = NO_EDGE;
   range->tree	   = RANGE;
   if (range->level < 3)
      return MAXCOSTS;
   if (range->x >= (unsigned) c->opt.width || range->y >= (unsigned) c->opt.height)
      return 0;

   /* motion compensation allowed for this range? (subdivide.c:141-147) */
   try_mc = (prediction && c->frame_type != 0
	     && range->level >= c->p_min_level && range->level <= c->p_max_level
	     && range->x + width_of_level (range->level) <= (unsigned) c->opt.width
	     && range->y + height_of_level (range->level) <= (unsigned) c->opt.height);
   try_nd = (prediction && c->frame_type == 0
	     && range->level >= c->p_min_level && range->level <= c->p_max_level);
   if (try_mc)
      clear_norms_table (range->level, c);

   if (range->level == (unsigned) c->opt.lc_max_level)
      init_range (range, band, c);

   price = c->price;
   if (band != 0)
      price *= c->opt.chroma_decrease;

   if (band != 0)
   {
      unsigned label;

      for (label = 0; label < MAXLABELS; label++)
	 if (y_state != RANGE)
	    new_y_state [label] = w->tree [y_state][label];
	 else
	    new_y_state [label] = RANGE;
   }
   else
      new_y_state [0] = new_y_state [1] = RANGE;

   /* snapshot of every model the recursion may modify (subdivide.c:188-194) */
   domain_model	     = malloc (sizeof (rle_model_t));
   lc_domain_model   = malloc (sizeof (rle_model_t));
   d_domain_model    = malloc (sizeof (rle_model_t));
   lc_d_domain_model = malloc (sizeof (rle_model_t));
   rle_copy (domain_model, &c->pool);
   rle_copy (d_domain_model, &c->d_pool);
   coeff_model	 = c->coeff.model;
   d_coeff_model = c->d_coeff.model;
   tree_model	 = c->tree;
   p_tree_model	 = c->p_tree;
   states	 = w->states;

   /* alternative 1: linear combination */
   if (range->level <= (unsigned) c->opt.lc_max_level)
   {
      lrange		     = *range;
      lrange.tree	     = RANGE;
      lrange.tree_bits	     = tree_bits (0, lrange.level, &c->tree);
      lrange.matrix_bits     = 0;
      lrange.weights_bits    = 0;
      lrange.mv_tree_bits    = try_mc ? 1 : 0;	/* mc allowed but not used */
      lrange.mv_coord_bits   = 0;
      lrange.nd_tree_bits    = 0;
      lrange.nd_weights_bits = 0;
      lrange.prediction	     = 0;
      c->ap = delta ? &c->d_pool : &c->pool;
      c->ac = delta ? &c->d_coeff : &c->coeff;
      lincomb_costs = approximate_range (max_costs, price, c->opt.max_elements,
					 y_state, &lrange, c);
      if (c->trace)
      {
	 int e;

	 fprintf (c->trace, "lc %u %u %u %u %u %u %d %u %08x %08x %08x",
		  c->lc_calls, lrange.level, lrange.image, lrange.address,
		  lrange.x, lrange.y, y_state, w->states, fbits (max_costs),
		  fbits (price), fbits (lincomb_costs));
	 if (lrange.into [0] != NO_EDGE)
	 {
	    fprintf (c->trace, " %08x %08x %08x :", fbits (lrange.err),
		     fbits (lrange.matrix_bits), fbits (lrange.weights_bits));
	    for (e = 0; lrange.into [e] != NO_EDGE; e++)
	       fprintf (c->trace, " %d:%08x", (int) lrange.into [e],
			fbits (lrange.weight [e]));
	 }
	 fprintf (c->trace, "\n");
      }
      c->lc_calls++;
   }
   else
      lincomb_costs = MAXCOSTS;

   /* keep the "lc" models, restore the snapshot (subdivide.c:226-237) */
   rle_copy (lc_domain_model, &c->pool);
   rle_copy (lc_d_domain_model, &c->d_pool);
   lc_coeff_model   = c->coeff.model;
   lc_d_coeff_model = c->d_coeff.model;
   rle_copy (&c->pool, domain_model);
   rle_copy (&c->d_pool, d_domain_model);
   c->coeff.model   = coeff_model;
   c->d_coeff.model = d_coeff_model;

   /* alternative 2: recursive subdivision */
   if (range->level > (unsigned) c->opt.lc_min_level)
   {
      unsigned label;

      memset (child, 0, sizeof child);
      rrange		     = *range;
      rrange.tree_bits	     = tree_bits (1, rrange.level, &c->tree);
      rrange.matrix_bits     = 0;
      rrange.weights_bits    = 0;
      rrange.err	     = 0;
      rrange.mv_tree_bits    = try_mc ? 1 : 0;
      rrange.mv_coord_bits   = 0;
      rrange.nd_tree_bits    = try_nd ? tree_bits (1, lrange.level, &c->p_tree) : 0;
      rrange.nd_weights_bits = 0;
      rrange.prediction	     = 0;
      subdivide_costs = (rrange.tree_bits + rrange.weights_bits
			 + rrange.matrix_bits + rrange.mv_tree_bits
			 + rrange.mv_coord_bits + rrange.nd_tree_bits
			 + rrange.nd_weights_bits) * price;

      for (label = 0; label < MAXLABELS; label++)
      {
	 float remaining_costs;

	 child [label].image   = rrange.image * MAXLABELS + label + 1;
	 child [label].address = rrange.address * MAXLABELS + label;
	 child [label].level   = rrange.level - 1;
	 child [label].x = rrange.level & 1
			   ? rrange.x
			   : rrange.x + label * width_of_level (rrange.level - 1);
	 child [label].y = rrange.level & 1
			   ? rrange.y + label * height_of_level (rrange.level - 1)
			   : rrange.y;

	 if (label && rrange.level <= (unsigned) c->opt.lc_max_level)
	    compute_ip_images_state (child [label].image, child [label].address,
				     child [label].level, 1, states, c);

	 remaining_costs = fmin2 (lincomb_costs, max_costs) - subdivide_costs;
	 if (remaining_costs > 0)
	    subdivide_costs += subdivide (remaining_costs, band, new_y_state [label],
					  &child [label], c, prediction, delta);
	 else if (try_mc && child [label].level >= c->p_min_level)
	    fill_norms_table (child [label].x, child [label].y, child [label].level, c);

	 if (try_mc)
	    update_norms_table (rrange.level, c);

	 if (subdivide_costs >= fmin2 (lincomb_costs, max_costs))
	 {
	    subdivide_costs = MAXCOSTS;
	    break;
	 }
	 rrange.err		+= child [label].err;
	 rrange.tree_bits	+= child [label].tree_bits;
	 rrange.matrix_bits	+= child [label].matrix_bits;
	 rrange.weights_bits	+= child [label].weights_bits;
	 rrange.mv_tree_bits	+= child [label].mv_tree_bits;
	 rrange.mv_coord_bits	+= child [label].mv_coord_bits;
	 rrange.nd_weights_bits += child [label].nd_weights_bits;
	 rrange.nd_tree_bits	+= child [label].nd_tree_bits;

	 tree_update (child [label].tree != RANGE, child [label].level, &c->tree);
	 tree_update (child [label].prediction ? 0 : 1, child [label].level, &c->p_tree);
      }
   }
   else
      subdivide_costs = MAXCOSTS;

   /* alternative 3: motion compensation + approximation of the prediction error
      (subdivide.c:383-407) */
   if (try_mc || try_nd)
   {
      const float prediction_costs
	 = predict_range (fmin2 (fmin2 (lincomb_costs, subdivide_costs), max_costs), price,
			  range, c, band, y_state, states, &tree_model, &p_tree_model,
			  domain_model, d_domain_model, &coeff_model, &d_coeff_model);

      if (prediction_costs < MAXCOSTS)
      {
	 free (domain_model);
	 free (lc_domain_model);
	 free (d_domain_model);
	 free (lc_d_domain_model);
	 return prediction_costs;
      }
   }

   if (lincomb_costs >= MAXCOSTS && subdivide_costs >= MAXCOSTS)
   {
      rle_copy (&c->pool, domain_model);
      rle_copy (&c->d_pool, d_domain_model);
      c->coeff.model   = coeff_model;
      c->d_coeff.model = d_coeff_model;
      c->tree	       = tree_model;
      c->p_tree	       = p_tree_model;
      if (w->states != states)
	 remove_states (states, w);
      subdivide_costs = MAXCOSTS;
   }
   else if (lincomb_costs < subdivide_costs)
   {
      rle_copy (&c->pool, lc_domain_model);
      rle_copy (&c->d_pool, lc_d_domain_model);
      c->coeff.model   = lc_coeff_model;
      c->d_coeff.model = lc_d_coeff_model;
      c->tree	       = tree_model;
      c->p_tree	       = p_tree_model;
      *range	       = lrange;
      if (w->states != states)
	 remove_states (states, w);
      subdivide_costs = lincomb_costs;
   }
   else
   {
      int aux = band > 0
		|| range->x + width_of_level (range->level) > (unsigned) c->opt.width
		|| range->y + height_of_level (range->level) > (unsigned) c->opt.height;

      init_new_state (aux, delta, &rrange, child, new_y_state, c);
      *range = rrange;
   }
   free (domain_model);
   free (lc_domain_model);
   free (d_domain_model);
   free (lc_d_domain_model);
   return subdivide_costs;
}

/*****************************************************************************
		  chroma pool  (domain-pool.c:854-879, wfalib.c:182-231)
*****************************************************************************/

typedef struct pair
{
   int16_t key, value;
} pair_t;

static int
cmp_desc_pair (const void *a, const void *b) /* lib/misc.c sort_desc_pair */
{
   return (int) ((const pair_t *) b)->key - (int) ((const pair_t *) a)->key;
}

static void
rle_chroma (unsigned max_domains, coder_t *c)
{
   const fo_wfa_t *w = c->wfa;
   rle_model_t	  *m = &c->pool;

   if (max_domains < m->n)
   {
      /* compute_hits (basis_states, states - 1, max_domains) */
      unsigned from = w->basis_states, to = w->states - 1, n = max_domains;
      unsigned state, label, edge;
      int      domain;
      pair_t  *hits    = calloc (to, sizeof (pair_t));
      int16_t *domains;

      for (domain = 0; domain < (int) to; domain++)
      {
	 hits [domain].value = (int16_t) domain;
	 hits [domain].key   = 0;
      }
      for (state = from; state <= to; state++)
	 for (label = 0; label < MAXLABELS; label++)
	    for (edge = 0; (domain = w->into [state][label][edge]) != NO_EDGE;
		 edge++)
	       hits [domain].key++;
      qsort (hits + 1, to - 1, sizeof (pair_t), cmp_desc_pair);
      n	      = fmin2 (to, n);
      domains = calloc (n + 1, sizeof (int16_t));
      for (domain = 0; domain < (int) n && (!domain || hits [domain].key);
	   domain++)
	 domains [domain] = hits [domain].value;
      n = (unsigned) domain;
      qsort (domains, n, sizeof (int16_t), cmp_word);
      domains [n] = -1;
      free (hits);

      for (n = 0; n < max_domains && domains [n] >= 0; n++)
	 m->states [n] = domains [n];
      max_domains = fmin2 (max_domains, n);
      free (domains);
      m->n = (uint16_t) max_domains;
   }
   m->y_index	  = 0;
   m->max_domains = m->n;
}

/*****************************************************************************
				public code
*****************************************************************************/

unsigned
fo_image_level (unsigned width, unsigned height) /* coder.c:249-256 */
{
   unsigned lx = (unsigned) (log2 ((double) (width - 1)) + 1);
   unsigned ly = (unsigned) (log2 ((double) (height - 1)) + 1);

   return (lx > ly ? lx : ly) * 2 - ((ly == lx + 1) ? 1 : 0);
}

void
fo_default_params (fo_params_t *p, int width, int height, int color,
		   float quality, int optimize)
/* CLI defaults: bin/cwfa.c:36-90 and :326-345 */
{
   memset (p, 0, sizeof *p);
   p->width		= width;
   p->height		= height;
   p->color		= color;
   p->quality		= quality;
   p->images_level	= 5;
   p->max_states	= 10000;
   p->chroma_max_states = 40;
   p->chroma_decrease	= 2.0f;
   p->rpf_mantissa	= 3;
   p->rpf_range_e	= 2;
   p->dc_rpf_mantissa	= 5;
   p->dc_rpf_range_e	= 1;
   if (optimize <= 0)
   {
      p->lc_min_level = 6;
      p->lc_max_level = 10;
      p->max_elements = 3;
      optimize	      = 0;
   }
   else
   {
      p->lc_min_level = 4;
      p->lc_max_level = 12;
      p->max_elements = 5;
      optimize	     -= 1;
   }
   p->second_domain_block = optimize > 0;
   p->check_for_overflow  = optimize > 1;
   p->check_for_underflow = optimize > 1;
   p->full_search	  = optimize > 1;
}

void
fo_grey_to_plane (const uint8_t *grey, size_t n, int16_t *plane) /* image.c:352-363 */
{
   size_t i;

   for (i = 0; i < n; i++)
      plane [i] = (int16_t) (((int) grey [i] - 128) * 16);
}

void
fo_rgb_to_planes (const uint8_t *rgb, size_t n, int16_t *y, int16_t *cb,
		  int16_t *cr) /* image.c:365-386 */
{
   size_t i;

   for (i = 0; i < n; i++)
   {
      int red = rgb [3 * i], green = rgb [3 * i + 1], blue = rgb [3 * i + 2];

      y [i]  = (int16_t) ((+0.2989 * red + 0.5866 * green + 0.1145 * blue - 128) * 16);
      cb [i] = (int16_t) ((-0.1687 * red - 0.3312 * green + 0.5000 * blue) * 16);
      cr [i] = (int16_t) ((+0.5000 * red - 0.4183 * green - 0.0816 * blue) * 16);
   }
}

int
fo_encode (const fo_params_t *p, const int16_t *const planes [3], fo_wfa_t *out,
	   fo_stats_t *stats, FILE *trace, char *errbuf, size_t errlen)
{
   static int  tables_ready = 0;
   coder_t    *c	    = calloc (1, sizeof (coder_t));
   fo_stats_t  dummy;
   fo_wfa_t   *w = out;
   unsigned    state, label, level;
   int	       rc = 0;

   if (!tables_ready)
   {
      init_matrix_probabilities ();
      tables_ready = 1;
   }
   memset (&dummy, 0, sizeof dummy);
   if (stats)
      memset (stats, 0, sizeof *stats);
   c->st     = stats ? stats : &dummy;
   c->trace  = trace;
   c->errbuf = errbuf;
   c->errlen = errlen;
   c->wfa    = w;
   c->opt    = *p;

   if (setjmp (c->env))
   {
      rc = 1;
      goto cleanup;
   }
   if ((p->width & 1) || (p->height & 1))
      fail (c, "Width and height of images must be even numbers.");
   if (p->quality <= 0)
      fail (c, "Compression quality has to be positive.");

   /* alloc_wfa (wfalib.c:45-121) */
   memset (w, 0, sizeof *w);
   for (state = 0; state < MAXSTATES; state++)
      for (label = 0; label < MAXLABELS; label++)
      {
	 w->into [state][label][0] = NO_EDGE;
	 w->tree [state][label]	   = RANGE;
	 w->y_state [state][label] = RANGE;
      }

   /* alloc_coder (coder.c:249-327) */
   c->level = w->level = fo_image_level ((unsigned) p->width, (unsigned) p->height);
   c->opt.lc_min_level = p->lc_min_level > 3 ? p->lc_min_level : 3;
   c->opt.lc_max_level = fmin2 (p->lc_max_level, (int) c->level - 1);
   /* tiling->exponent is always 0 (quirk C1), so coder.c:273-279 reduces to: */
   if (c->opt.lc_max_level >= (int) c->level)
      c->opt.lc_max_level = (int) c->level - 1;
   if (c->opt.lc_min_level > c->opt.lc_max_level)
      c->opt.lc_min_level = c->opt.lc_max_level;
   c->opt.images_level = fmin2 (p->images_level, c->opt.lc_max_level - 1);
   c->products_level
      = (unsigned) (c->opt.lc_max_level - c->opt.images_level - 1 > 0
		    ? c->opt.lc_max_level - c->opt.images_level - 1 : 0);
   c->pixels = calloc (size_of_level ((unsigned) c->opt.lc_max_level),
		       sizeof (float));
   {
      int ms = fmin2 (p->max_states, MAXSTATES);
      c->opt.max_states = ms > 1 ? ms : 1;
      ms = fmin2 (p->max_elements, MAXEDGES);
      c->opt.max_elements = ms > 1 ? ms : 1;
      c->opt.chroma_max_states = p->chroma_max_states > 1 ? p->chroma_max_states : 1;
   }
   c->rpf    = make_rpf ((unsigned) p->rpf_mantissa, p->rpf_range_e);
   c->dc_rpf = make_rpf ((unsigned) p->dc_rpf_mantissa, p->dc_rpf_range_e);
   for (state = 0; state < 3; state++)
      c->planes [state] = planes [state];

   append_basis_states (c);			/* coder.c:161-162 */
   c->price = 128 * 64 / p->quality;		/* coder.c:164 */

   /* frame_coder (coder.c:692-892) */
   init_tree_model (&c->tree);
   rle_init (&c->pool, (unsigned) c->opt.max_states);
   for (state = 0; state < w->basis_states; state++)
      if (usedomain (state, w))
	 rle_append (&c->pool, state);
   aac_init (&c->coeff, c->rpf, c->dc_rpf, (unsigned) c->opt.lc_min_level,
	     (unsigned) c->opt.lc_max_level);
   /* the delta side of a still: "constant" pool, its own (never used) coefficient model */
   init_tree_model (&c->p_tree);
   rle_init (&c->d_pool, (unsigned) c->opt.max_states);
   aac_init (&c->d_coeff, c->rpf, c->dc_rpf, (unsigned) c->opt.lc_min_level,
	     (unsigned) c->opt.lc_max_level);
   c->d_pool_is_rle = 0;
   c->frame_type    = 0;
   c->ap	    = &c->pool;
   c->ac	    = &c->coeff;

   if (!p->color)
   {
      range_t range;

      memset (&range, 0, sizeof range);
      range.level = c->level;
      out->costs [0] = subdivide (MAXCOSTS, 0, RANGE, &range, c, 0, 0);
      if (range.tree == RANGE)
	 fail (c, "No root state generated!");
      w->root_state	    = (unsigned) range.tree;
      out->err [0]	    = range.err;
      out->tree_bits [0]    = range.tree_bits;
      out->matrix_bits [0]  = range.matrix_bits;
      out->weights_bits [0] = range.weights_bits;
   }
   else
   {
      int      YCb_node = -1, tree [3];
      unsigned band;

      for (band = 0; band < 3; band++)
      {
	 range_t range;

	 tree [band] = RANGE;
	 if (band == 1)
	 {
	    unsigned min_level;

	    rle_chroma ((unsigned) c->opt.chroma_max_states, c);
	    for (min_level = MAXLEVEL, state = w->basis_states;
		 state < w->states; state++)
	    {
	       unsigned lincomb = 0;

	       for (label = 0; label < MAXLABELS; label++)
		  lincomb += w->tree [state][label] == RANGE ? 1 : 0;
	       if (lincomb)
		  min_level = fmin2 (min_level,
				     (unsigned) (w->level_of_state [state] - 1));
	    }
	    c->opt.lc_min_level = (int) min_level;
	 }
	 memset (&range, 0, sizeof range);
	 range.level = c->level;
	 out->costs [band] = subdivide (MAXCOSTS, band, tree [0], &range, c, 0, 0);
	 out->err [band]	  = range.err;
	 out->tree_bits [band]	  = range.tree_bits;
	 out->matrix_bits [band]  = range.matrix_bits;
	 out->weights_bits [band] = range.weights_bits;
	 if (range.tree == RANGE)
	    fail (c, "No root state generated for color component %d!", band);
	 tree [band] = range.tree;
	 if (band == 1)
	 {
	    w->tree [w->states][0] = (int16_t) tree [0];
	    w->tree [w->states][1] = (int16_t) tree [1];
	    YCb_node		   = (int) w->states;
	    append_state (1, compute_final_distribution (w->states, w),
			  c->level + 1, c);
	 }
      }
      w->tree [w->states][0] = (int16_t) tree [2];
      w->tree [w->states][1] = RANGE;
      append_state (1, compute_final_distribution (w->states, w), c->level + 1, c);
      w->tree [w->states][0] = (int16_t) YCb_node;
      w->tree [w->states][1] = (int16_t) (w->states - 1);
      append_state (1, compute_final_distribution (w->states, w), c->level + 2, c);
      w->root_state = w->states - 1;
   }

cleanup:
   for (state = 0; state < MAXSTATES; state++)
   {
      free (c->images_of_state [state]);
      free (c->ip_images_state [state]);
      for (level = 0; level < MAXLEVEL; level++)
	 free (c->ip_states_state [state][level]);
   }
   free (c->pixels);
   free (c);
   return rc;
}

void
fo_dump_wfa (const fo_wfa_t *w, const fo_params_t *p, FILE *f)
{
   unsigned state, label, edge;

   (void) p;
   fprintf (f, "frame 0 0 %u %u\n", w->states, w->root_state);
   for (state = w->basis_states; state < w->states; state++)
   {
      fprintf (f, "s %u %d %d %d %u %u %u %u 0 0\n", state,
	       (int) w->level_of_state [state], (int) w->tree [state][0],
	       (int) w->tree [state][1], (unsigned) w->x [state][0],
	       (unsigned) w->y [state][0], (unsigned) w->x [state][1],
	       (unsigned) w->y [state][1]);
      for (label = 0; label < MAXLABELS; label++)
	 for (edge = 0; w->into [state][label][edge] != NO_EDGE; edge++)
	    fprintf (f, "e %u %u %d %08x %.9g\n", state, label,
		     (int) w->into [state][label][edge],
		     fbits (w->weight [state][label][edge]),
		     (double) w->weight [state][label][edge]);
   }
   fprintf (f, "end\n");
}

/*****************************************************************************

	    image regeneration from an automaton  (codec/decoder.c)

  The coder regenerates every frame it has coded (codec/coder.c:647-651): the result is
  the reference frame of the next predicted frame, so this side of the codec belongs to
  the encoder path of videos.  decode_image (decoder.c:412-535) with
  alloc_state_images (:878-1016) and compute_state_images (:1107-1498), restated per
  pixel: the reference adds two pixels per 32-bit int with a guard bit each (bit 0 and
  bit 16 are masked after every addition), which is a wrapping 16-bit addition of even
  numbers per pixel.

*****************************************************************************/

typedef struct simg
{
   int16_t *pix [MAXSTATES][MAXLEVEL + 1];	/* image of a state at a level, or NULL */
   const fo_wfa_t *w;
} simg_t;

/* codec/wfalib.c:273: the integer form of a weight the decoder multiplies with */
static int
int_weight_of (float weight)
{
   return (int16_t) (weight * 512 + 0.5);
}

/* image of 'state' at 'level' (width_of_level x height_of_level shorts, row-major) */
static const int16_t *
state_image (simg_t *si, unsigned state, unsigned level)
{
   const fo_wfa_t *w = si->w;

   if (si->pix [state][level])
      return si->pix [state][level];

   int16_t *img = calloc (size_of_level (level), sizeof (int16_t));

   si->pix [state][level] = img;
   if (level == 0)			/* decoder.c:1128-1130 */
   {
      img [0] = (int16_t) ((int) (w->final_distribution [state] * 8 + .5) * 2);
      return img;
   }
   const unsigned width	 = width_of_level (level - 1);
   const unsigned height = height_of_level (level - 1);
   const unsigned stride = width_of_level (level);

   for (unsigned label = 0; label < MAXLABELS; label++)
   {
      /* odd levels are split into an upper and a lower half, even ones into a left and a
	 right half (decoder.c:1166-1177) */
      int16_t *range = (level & 1) ? img + label * height * stride : img + label * width;
      int      child = w->tree [state][label];

      if (child != RANGE)		/* copy the child's image (decoder.c:1197-1209) */
      {
	 const int16_t *src = state_image (si, (unsigned) child, level - 1);

	 for (unsigned y = 0; y < height; y++)
	    memcpy (range + y * stride, src + y * width, width * sizeof (int16_t));
      }
      for (unsigned edge = 0; w->into [state][label][edge] != NO_EDGE; edge++)
      {
	 const int domain = w->into [state][label][edge];

	 if (domain != 0)		/* decoder.c:1219-1300, 1355-1440 */
	 {
	    const int16_t *src	  = state_image (si, (unsigned) domain, level - 1);
	    const int	   weight = int_weight_of (w->weight [state][label][edge]);

	    for (unsigned y = 0; y < height; y++)
	       for (unsigned x = 0; x < width; x++)
	       {
		  const int v = ((weight * (int) src [y * width + x]) >> 10) << 1;

		  range [y * stride + x] = (int16_t) (range [y * stride + x] + v);
	       }
	 }
	 else				/* the constant state (decoder.c:1302-1345, 1442-1492) */
	 {
	    const int weight = (int) (w->weight [state][label][edge]
				      * w->final_distribution [0] * 8 + .5) * 2;

	    for (unsigned y = 0; y < height; y++)
	       for (unsigned x = 0; x < width; x++)
		  range [y * stride + x] = (int16_t) (range [y * stride + x] + weight);
	 }
      }
   }
   return img;
}

/*
 *  decode_image (decoder.c:412-535), 4:4:4.  planes [b] receive width * height shorts each
 *  (1 band grey, 3 bands Y Cb Cr).  Returns 0 on success.
 */
int
fo_decode_image (const fo_wfa_t *w, int color, unsigned width, unsigned height,
		 int16_t *const planes [3])
{
   simg_t  *si = calloc (1, sizeof *si);
   unsigned root [3], max_level = 0, aw = 0, ah = 0, state;
   uint8_t  level_of_state [MAXSTATES];

   if (!si)
      return 1;
   si->w = w;
   memcpy (level_of_state, w->level_of_state, sizeof level_of_state);
   if (color)				/* decoder.c:436-444, 467-472 */
   {
      root [0] = (unsigned) w->tree [w->tree [w->root_state][0]][0];
      root [1] = (unsigned) w->tree [w->tree [w->root_state][0]][1];
      root [2] = (unsigned) w->tree [w->tree [w->root_state][1]][0];
      level_of_state [w->root_state]		       = 128;
      level_of_state [w->tree [w->root_state][0]] = 128;
      level_of_state [w->tree [w->root_state][1]] = 128;
   }
   else
      root [0] = root [1] = root [2] = w->root_state;
   /* highest level of a linear combination; frame size the bintree covers
      (decoder.c:449-461, compute_actual_size :843-875 for 4:4:4) */
   for (state = w->basis_states; state < w->states; state++)
      if (w->into [state][0][0] != NO_EDGE || w->into [state][1][0] != NO_EDGE)
      {
	 const unsigned l = w->level_of_state [state];

	 if (l > max_level)
	    max_level = l;
	 if (w->x [state][0] + width_of_level (l) > aw)
	    aw = w->x [state][0] + width_of_level (l);
	 if (w->y [state][0] + height_of_level (l) > ah)
	    ah = w->y [state][0] + height_of_level (l);
      }
   aw += aw & 1;
   ah += ah & 1;
   if (aw < width)
      aw = width;
   if (ah < height)
      ah = height;

   const unsigned bands = color ? 3 : 1;
   int16_t	 *frame [3] = {NULL, NULL, NULL};

   for (unsigned b = 0; b < bands; b++)
      frame [b] = calloc ((size_t) aw * ah, sizeof (int16_t));
   /* every state of level max_level is one block of the frame (decoder.c:913-937) */
   for (state = w->basis_states; state < w->states; state++)
      if (level_of_state [state] == max_level)
      {
	 const unsigned b   = !color || state <= root [0] ? 0 : state > root [1] ? 2 : 1;
	 const unsigned bw  = width_of_level (max_level), bh = height_of_level (max_level);
	 const int16_t *img = state_image (si, state, max_level);
	 int16_t       *dst = frame [b] + (size_t) w->y [state][0] * aw + w->x [state][0];

	 /* the block may reach beyond the frame where the bintree has no (visible) ranges:
	    the reference computes such blocks in place and simply never writes there */
	 const unsigned cw = w->x [state][0] >= aw ? 0 : (aw - w->x [state][0] < bw ? aw - w->x [state][0] : bw);

	 for (unsigned y = 0; y < bh && w->y [state][0] + y < ah; y++)
	    memcpy (dst + (size_t) y * aw, img + y * bw, cw * sizeof (int16_t));
      }
   /* crop to the size at coding time (decoder.c:500-527) */
   for (unsigned b = 0; b < bands; b++)
   {
      for (unsigned y = 0; y < height; y++)
	 memcpy (planes [b] + (size_t) y * width, frame [b] + (size_t) y * aw,
		 width * sizeof (int16_t));
      free (frame [b]);
   }
   for (state = 0; state < MAXSTATES; state++)
      for (unsigned l = 0; l <= MAXLEVEL; l++)
	 free (si->pix [state][l]);
   free (si);
   return 0;
}

/*****************************************************************************

	     motion compensation of a regenerated frame  (codec/motion.c)

*****************************************************************************/

/* extract_mc_block (motion.c:232-334): the reference block displaced by (mx, my) */
static void
extract_mc_block (int16_t *mcblock, unsigned width, unsigned height,
		  const int16_t *reference, unsigned ref_width, int half_pixel,
		  unsigned xo, unsigned yo, int mx, int my)
{
   if (!half_pixel)
   {
      /* the reference adds the vector as unsigned numbers: same result modulo 2^32 */
      const int16_t *rblock = reference + (ptrdiff_t) ((int) yo + my) * ref_width + ((int) xo + mx);

      for (unsigned y = 0; y < height; y++)
	 memcpy (mcblock + y * width, rblock + (size_t) y * ref_width, width * sizeof (int16_t));
   }
   else
   {
      const int16_t *r = reference + (ptrdiff_t) ((int) yo + my / 2) * ref_width + ((int) xo + mx / 2);

      for (unsigned y = 0; y < height; y++)
	 for (unsigned x = 0; x < width; x++)
	 {
	    const int16_t *q = r + (size_t) y * ref_width + x;

	    if (!(mx & 1) && !(my & 1))
	       mcblock [y * width + x] = q [0];
	    else if (!(mx & 1))
	       mcblock [y * width + x] = (int16_t) ((q [0] + q [ref_width]) >> 1);
	    else if (!(my & 1))
	       mcblock [y * width + x] = (int16_t) ((q [0] + q [1]) >> 1);
	    else
	       mcblock [y * width + x] = (int16_t) ((q [0] + q [1] + q [ref_width] + q [ref_width + 1]) >> 2);
	 }
   }
}

static void
restore_mc_band (const fo_wfa_t *w, unsigned root_state, unsigned width, int half_pixel,
		 int16_t *image, const int16_t *past, const int16_t *future)
{
   int16_t *mcblock  = malloc (size_of_level (MAXLEVEL > 16 ? 16 : MAXLEVEL) * sizeof (int16_t));
   int16_t *mcblock2 = malloc (size_of_level (MAXLEVEL > 16 ? 16 : MAXLEVEL) * sizeof (int16_t));

   for (unsigned state = w->basis_states; state <= root_state; state++)
      for (unsigned label = 0; label < MAXLABELS; label++)
	 if (w->mv_type [state][label] != 0)	/* motion.c:69-190 */
	 {
	    const unsigned level = w->level_of_state [state] - 1u;
	    const unsigned bw = width_of_level (level), bh = height_of_level (level);
	    const int	   type	 = w->mv_type [state][label];

	    if (type == 1 || type == 3)
	       extract_mc_block (mcblock, bw, bh, past, width, half_pixel, w->x [state][label],
				 w->y [state][label], w->mv_fx [state][label], w->mv_fy [state][label]);
	    if (type == 2 || type == 3)
	       extract_mc_block (type == 2 ? mcblock : mcblock2, bw, bh, future, width, half_pixel,
				 w->x [state][label], w->y [state][label], w->mv_bx [state][label],
				 w->mv_by [state][label]);
	    for (unsigned y = 0; y < bh; y++)
	       for (unsigned x = 0; x < bw; x++)
	       {
		  int16_t  *o	= image + (size_t) (w->y [state][label] + y) * width + w->x [state][label] + x;
		  const int ref = type == 3 ? (mcblock [y * bw + x] + mcblock2 [y * bw + x]) >> 1
					    : mcblock [y * bw + x];

		  *o = (int16_t) (*o + ref);
	       }
	 }
   free (mcblock);
   free (mcblock2);
}

void
fo_restore_mc (const fo_wfa_t *w, unsigned width, unsigned height, int half_pixel,
	       int16_t *image, const int16_t *past, const int16_t *future)
{
   (void) height;
   restore_mc_band (w, w->root_state, width, half_pixel, image, past, future);
}

/* restore_mc for a colour frame (4:4:4): the vectors of the luminance tree -- the states up to the
   root of the Y band -- move all three bands, then the chroma bands are clipped to 8 bits
   (motion.c:59-62, 192-224).  image / past / future: three planes of width * height shorts. */
void
fo_restore_mc_colour (const fo_wfa_t *w, unsigned width, unsigned height, int16_t *image,
		      const int16_t *past, const int16_t *future)
{
   const size_t	  npix	 = (size_t) width * height;
   const unsigned y_root = (unsigned) w->tree [w->tree [w->root_state][0]][0];

   for (unsigned b = 0; b < 3; b++)
      restore_mc_band (w, y_root, width, 0, image + b * npix, past ? past + b * npix : NULL,
		       future ? future + b * npix : NULL);
   for (size_t n = npix; n < 3 * npix; n++)
   {
      int v = image [n] >> 4;		/* HAVE_SIGNED_SHIFT (oracle/refcfg/config.h) */

      v = v < -128 ? -128 : v > 127 ? 127 : v;
      image [n] = (int16_t) (v * 16);
   }
}

/* subtract_mc (mwfa.c:156-299): before the chroma bands of a predicted colour frame are coded, the
   motion compensation of the luminance tree is taken off the chroma planes of the original -- with
   the vector components rounded towards zero to even numbers ((v / 2) * 2), which restore_mc does
   NOT do: reproduced as it is. */
static void
subtract_mc (const fo_wfa_t *w, unsigned width, int16_t *const chroma [2],
	     const int16_t *const past [2], const int16_t *const future [2])
{
   int16_t *mcblock  = malloc (size_of_level (MAXLEVEL > 16 ? 16 : MAXLEVEL) * sizeof (int16_t));
   int16_t *mcblock2 = malloc (size_of_level (MAXLEVEL > 16 ? 16 : MAXLEVEL) * sizeof (int16_t));

   for (unsigned state = w->basis_states; state < w->states; state++)
      for (unsigned label = 0; label < MAXLABELS; label++)
	 if (w->mv_type [state][label] != 0)
	 {
	    const unsigned level = w->level_of_state [state] - 1u;
	    const unsigned bw = width_of_level (level), bh = height_of_level (level);
	    const int	   type	 = w->mv_type [state][label];

	    for (unsigned b = 0; b < 2; b++)
	    {
	       if (type == 1 || type == 3)
		  extract_mc_block (mcblock, bw, bh, past [b], width, 0, w->x [state][label],
				    w->y [state][label], (w->mv_fx [state][label] / 2) * 2,
				    (w->mv_fy [state][label] / 2) * 2);
	       if (type == 2 || type == 3)
		  extract_mc_block (type == 2 ? mcblock : mcblock2, bw, bh, future [b], width, 0,
				    w->x [state][label], w->y [state][label],
				    (w->mv_bx [state][label] / 2) * 2, (w->mv_by [state][label] / 2) * 2);
	       for (unsigned y = 0; y < bh; y++)
		  for (unsigned x = 0; x < bw; x++)
		  {
		     int16_t  *o   = chroma [b] + (size_t) (w->y [state][label] + y) * width
				     + w->x [state][label] + x;
		     /* (a division here, mwfa.c:287, a shift in restore_mc) */
		     const int ref = type == 3 ? (mcblock [y * bw + x] + mcblock2 [y * bw + x]) / 2
					       : mcblock [y * bw + x];

		     *o = (int16_t) (*o - ref);
		  }
	    }
	 }
   free (mcblock);
   free (mcblock2);
}

int
fo_wfa_from_dump (const char *text, unsigned root_state, fo_wfa_t *w)
{
   memset (w, 0, sizeof *w);
   for (unsigned s = 0; s < MAXSTATES; s++)
      for (unsigned l = 0; l < MAXLABELS; l++)
      {
	 w->into [s][l][0] = NO_EDGE;
	 w->tree [s][l]	   = RANGE;
	 w->y_state [s][l] = RANGE;
      }
   /* input/basis.c:126-131 */
   w->basis_states = w->states = 3;
   w->final_distribution [0] = 128;
   w->final_distribution [1] = 64;
   w->final_distribution [2] = 64;
   append_edge (0, 0, 1.0f, 0, w);
   append_edge (0, 0, 1.0f, 1, w);
   append_edge (1, 2, 0.5f, 0, w);
   append_edge (1, 2, 0.5f, 1, w);
   append_edge (1, 0, 0.5f, 1, w);
   append_edge (2, 1, 1.0f, 0, w);
   append_edge (2, 1, 1.0f, 1, w);
   for (unsigned s = 0; s < 3; s++)
      w->level_of_state [s] = (uint8_t) -1;

   const char *p = text;

   while (*p)
   {
      unsigned s, l, a, b, c2, d2;
      int      t0, t1, lv, into, ty, fx, fy, bx, by;
      unsigned bits;

      if (sscanf (p, "s %u %d %d %d %u %u %u %u", &s, &lv, &t0, &t1, &a, &b, &c2, &d2) == 8)
      {
	 if (s >= MAXSTATES)
	    return 1;
	 w->level_of_state [s] = (uint8_t) lv;
	 w->tree [s][0] = (int16_t) t0;
	 w->tree [s][1] = (int16_t) t1;
	 w->x [s][0] = (uint16_t) a;
	 w->y [s][0] = (uint16_t) b;
	 w->x [s][1] = (uint16_t) c2;
	 w->y [s][1] = (uint16_t) d2;
	 if (s + 1 > w->states)
	    w->states = s + 1;
      }
      else if (sscanf (p, "e %u %u %d %x", &s, &l, &into, &bits) == 4)
      {
	 float    f;
	 unsigned e;

	 memcpy (&f, &bits, sizeof f);
	 for (e = 0; w->into [s][l][e] != NO_EDGE; e++)
	    ;
	 w->into [s][l][e]     = (int16_t) into;	/* the dump lists them in stored order */
	 w->weight [s][l][e]   = f;
	 w->into [s][l][e + 1] = NO_EDGE;
      }
      else if (sscanf (p, "m %u %u %d %d %d %d %d", &s, &l, &ty, &fx, &fy, &bx, &by) == 7)
      {
	 w->mv_type [s][l] = (int8_t) ty;
	 w->mv_fx [s][l]   = (int8_t) fx;
	 w->mv_fy [s][l]   = (int8_t) fy;
	 w->mv_bx [s][l]   = (int8_t) bx;
	 w->mv_by [s][l]   = (int8_t) by;
      }
      else if (sscanf (p, "d %u", &s) == 1)
	 w->delta_state [s] = 1;
      while (*p && *p != '\n')
	 p++;
      if (*p)
	 p++;
   }
   w->root_state = root_state;
   for (unsigned s = w->basis_states; s < w->states; s++)
      w->final_distribution [s] = compute_final_distribution (s, w);
   return 0;
}

/*****************************************************************************

	      sequences  (video_coder / frame_coder, codec/coder.c:490-892)

*****************************************************************************/

/*
 *  Encode a grey sequence: frame 0 intra, the others by 'pattern' (I or P per frame,
 *  coder.c:522-535, 684-706), every predicted frame against the regenerated previous
 *  frame (coder.c:560-651).  out [f] receives the automaton of frame f, reconst (if not
 *  NULL) the regenerated frames, width * height shorts each.  Returns 0 on success.
 */
int
fo_encode_video (const fo_params_t *p, int n_frames, const int16_t *const *frames,
		 const char *pattern, int p_min_level, int p_max_level, int search_range,
		 fo_wfa_t *out, int16_t *reconst, FILE *trace, char *errbuf, size_t errlen)
{
   static int  tables_ready = 0;
   coder_t    *c = calloc (1, sizeof (coder_t));
   fo_wfa_t   *w = calloc (1, sizeof (fo_wfa_t));
   fo_stats_t  dummy;
   const size_t npix = (size_t) p->width * p->height;
   const unsigned bands = p->color ? 3 : 1;	/* colour: frames [3 f + b], reconst 3 planes per frame */
   int16_t    *past = calloc (bands * npix, sizeof (int16_t)), *cur = calloc (bands * npix, sizeof (int16_t));
   int16_t    *chroma = p->color ? calloc (2 * npix, sizeof (int16_t)) : NULL;
   unsigned    state, label, level;
   int	       rc = 0;

   if (!tables_ready)
   {
      init_matrix_probabilities ();
      tables_ready = 1;
   }
   memset (&dummy, 0, sizeof dummy);
   c->st     = &dummy;
   c->trace  = trace;
   c->errbuf = errbuf;
   c->errlen = errlen;
   c->wfa    = w;
   c->opt    = *p;
   if (setjmp (c->env))
   {
      rc = 1;
      goto cleanup;
   }
   if ((p->width & 1) || (p->height & 1))
      fail (c, "Width and height of images must be even numbers.");
   if (p->quality <= 0)
      fail (c, "Compression quality has to be positive.");
   if (search_range < 1 || search_range > 16)
      fail (c, "search range must be in 1..16");

   for (state = 0; state < MAXSTATES; state++)
      for (label = 0; label < MAXLABELS; label++)
      {
	 w->into [state][label][0] = NO_EDGE;
	 w->tree [state][label]	   = RANGE;
	 w->y_state [state][label] = RANGE;
      }
   /* alloc_coder (coder.c:249-327) */
   c->level = w->level = fo_image_level ((unsigned) p->width, (unsigned) p->height);
   c->opt.lc_min_level = p->lc_min_level > 3 ? p->lc_min_level : 3;
   c->opt.lc_max_level = fmin2 (p->lc_max_level, (int) c->level - 1);
   if (c->opt.lc_min_level > c->opt.lc_max_level)
      c->opt.lc_min_level = c->opt.lc_max_level;
   c->opt.images_level = fmin2 (p->images_level, c->opt.lc_max_level - 1);
   c->products_level
      = (unsigned) (c->opt.lc_max_level - c->opt.images_level - 1 > 0
		    ? c->opt.lc_max_level - c->opt.images_level - 1 : 0);
   c->pixels = calloc (size_of_level ((unsigned) c->opt.lc_max_level), sizeof (float));
   {
      int ms = fmin2 (p->max_states, MAXSTATES);
      c->opt.max_states = ms > 1 ? ms : 1;
      ms = fmin2 (p->max_elements, MAXEDGES);
      c->opt.max_elements = ms > 1 ? ms : 1;
      c->opt.chroma_max_states = p->chroma_max_states > 1 ? p->chroma_max_states : 1;
   }
   c->rpf    = make_rpf ((unsigned) p->rpf_mantissa, p->rpf_range_e);
   c->dc_rpf = make_rpf ((unsigned) p->dc_rpf_mantissa, p->dc_rpf_range_e);
   /* prediction levels are a subset of the range levels (coder.c:284-290) */
   c->p_min_level = (unsigned) (p_min_level > c->opt.lc_min_level ? p_min_level : c->opt.lc_min_level);
   c->p_max_level = (unsigned) fmin2 (p_max_level, c->opt.lc_max_level);
   if (c->p_min_level > c->p_max_level)
      c->p_min_level = c->p_max_level;
   c->search_range = (unsigned) search_range;
   /* alloc_motion (mwfa.c:86-126) */
   for (int dx = -search_range; dx < search_range; dx++)
      c->xbits [dx + search_range] = c->ybits [dx + search_range]
				   = (float) mv_code_length [dx + search_range];
   for (level = c->p_min_level; level <= c->p_max_level; level++)
   {
      c->mc_forward_norms [level]  = calloc (norms_size (c), sizeof (float));
      c->mc_backward_norms [level] = calloc (norms_size (c), sizeof (float));
   }

   append_basis_states (c);
   c->price = 128 * 64 / p->quality;

   /* video_coder (coder.c:490-680): frames are coded in display order except that a B frame
      waits for the next non-B frame (its future reference), which is coded first; the last
      frame of the sequence is forced to be a P frame */
   {
      int display = 0, future_display = -1, future_frame = 0, coded = 0;
      int16_t *fut = calloc (bands * npix, sizeof (int16_t));
      int have_reconst = 0;

      while (display < n_frames)
      {
	 range_t range;
	 int	 type, frame;

	 if (display == 0)
	    type = 0;
	 else
	 {
	    const int t = pattern [(size_t) display % strlen (pattern)];

	    type = (t == 'p' || t == 'P') ? 1 : (t == 'b' || t == 'B') ? 2 : (t == 'i' || t == 'I') ? 0 : -1;
	    if (type < 0)
	       fail (c, "Frame type %c not valid. Choose one of I,B or P.", t);
	 }
	 if (display == future_display)		/* already coded as a future reference */
	 {
	    display++;
	    continue;
	 }
	 else if (type == 2 && display > future_display)
	 {
	    int i = display;

	    while (type == 2)
	    {
	       i++;
	       if (i >= n_frames)
	       {
		  future_display = i - 1;
		  type		 = 1;
	       }
	       else
	       {
		  const int t = pattern [(size_t) i % strlen (pattern)];

		  future_display = i;
		  type = (t == 'p' || t == 'P') ? 1 : (t == 'b' || t == 'B') ? 2 : 0;
	       }
	       frame = future_display;
	    }
	 }
	 else
	 {
	    frame = display;
	    display++;
	 }

	 /* reference frames (coder.c:571-627); 'cur' holds the last regenerated frame */
	 if (type == 0)
	    have_reconst = 0;
	 else if (type == 1)
	 {
	    int16_t *t = past; past = cur; cur = t;	/* past <- current */
	    have_reconst = 0;
	 }
	 else if (future_frame)
	 {
	    int16_t *t = fut; fut = cur; cur = t;	/* future <- current */
	    have_reconst = 0;
	 }
	 else					/* B_as_past_ref = YES (options.c:99) */
	 {
	    int16_t *t = past; past = cur; cur = t;
	    have_reconst = 0;
	 }
	 (void) have_reconst;
	 future_frame  = frame == future_display;
	 c->frame_type = type;
	 c->planes [0] = frames [(size_t) frame * bands];
	 c->past       = past;
	 c->future     = fut;
	 if (p->color)		/* subtract_mc changes the chroma planes of the original */
	 {
	    memcpy (chroma, frames [(size_t) frame * 3 + 1], npix * sizeof (int16_t));
	    memcpy (chroma + npix, frames [(size_t) frame * 3 + 2], npix * sizeof (int16_t));
	    c->planes [1] = chroma;
	    c->planes [2] = chroma + npix;
	 }

	 /* frame_coder (coder.c:692-755) */
	 init_tree_model (&c->tree);
	 init_tree_model (&c->p_tree);
	 rle_init (&c->pool, (unsigned) c->opt.max_states);
	 rle_init (&c->d_pool, (unsigned) c->opt.max_states);
	 c->d_pool_is_rle = type != 0 || nd_prediction_on;	/* coder.c:720-725 */
	 for (state = 0; state < w->basis_states; state++)
	    if (usedomain (state, w))
	    {
	       rle_append (&c->pool, state);
	       if (c->d_pool_is_rle)
		  rle_append (&c->d_pool, state);
	    }
	 aac_init (&c->coeff, c->rpf, c->dc_rpf, (unsigned) c->opt.lc_min_level,
		   (unsigned) c->opt.lc_max_level);
	 aac_init (&c->d_coeff, c->rpf, c->dc_rpf, (unsigned) c->opt.lc_min_level,
		   (unsigned) c->opt.lc_max_level);
	 c->ap = &c->pool;
	 c->ac = &c->coeff;

	 w->frame_type	 = type;
	 w->frame_number = frame;
	 if (!p->color)
	 {
	 memset (&range, 0, sizeof range);
	 range.level = c->level;
	 w->costs [0] = subdivide (MAXCOSTS, 0, RANGE, &range, c, type != 0 || nd_prediction_on, 0);
	 if (range.tree == RANGE)
	    fail (c, "No root state generated!");
	 w->root_state	     = (unsigned) range.tree;
	 w->err [0]	     = range.err;
	 w->tree_bits [0]    = range.tree_bits;
	 w->matrix_bits [0]  = range.matrix_bits;
	 w->weights_bits [0] = range.weights_bits;
	 }
	 else
	 {
	    /* the colour bands (coder.c:757-849): Y with the frame's prediction, then -- the pool cut
	       down to the most used luminance states, the range levels limited to those the luminance
	       used (for good: the option is never set back), the luminance tree's motion compensation
	       taken off the chroma planes -- Cb and Cr without prediction; virtual states on top */
	    int	     YCb_node = -1, tree [3];
	    unsigned band;

	    for (band = 0; band < 3; band++)
	    {
	       tree [band] = RANGE;
	       if (band == 1)
	       {
		  unsigned min_level;

		  rle_chroma ((unsigned) c->opt.chroma_max_states, c);
		  for (min_level = MAXLEVEL, state = w->basis_states; state < w->states; state++)
		  {
		     unsigned lincomb = 0;

		     for (label = 0; label < MAXLABELS; label++)
			lincomb += w->tree [state][label] == RANGE ? 1 : 0;
		     if (lincomb)
			min_level = fmin2 (min_level, (unsigned) (w->level_of_state [state] - 1));
		  }
		  c->opt.lc_min_level = (int) min_level;
		  if (type != 0)
		  {
		     int16_t *const	  ch [2] = {chroma, chroma + npix};
		     const int16_t *const pp [2] = {past + npix, past + 2 * npix};
		     const int16_t *const ff [2] = {fut + npix, fut + 2 * npix};

		     subtract_mc (w, (unsigned) p->width, ch, pp, ff);
		  }
	       }
	       memset (&range, 0, sizeof range);
	       range.level = c->level;
	       w->costs [band] = subdivide (MAXCOSTS, band, tree [0], &range, c, type != 0 && band == 0, 0);
	       w->err [band]	      = range.err;
	       w->tree_bits [band]    = range.tree_bits;
	       w->matrix_bits [band]  = range.matrix_bits;
	       w->weights_bits [band] = range.weights_bits;
	       if (range.tree == RANGE)
		  fail (c, "No root state generated for color component %d!", band);
	       tree [band] = range.tree;
	       if (band == 1)
	       {
		  w->tree [w->states][0] = (int16_t) tree [0];
		  w->tree [w->states][1] = (int16_t) tree [1];
		  YCb_node		 = (int) w->states;
		  append_state (1, compute_final_distribution (w->states, w), c->level + 1, c);
	       }
	    }
	    w->tree [w->states][0] = (int16_t) tree [2];
	    w->tree [w->states][1] = RANGE;
	    append_state (1, compute_final_distribution (w->states, w), c->level + 1, c);
	    w->tree [w->states][0] = (int16_t) YCb_node;
	    w->tree [w->states][1] = (int16_t) (w->states - 1);
	    append_state (1, compute_final_distribution (w->states, w), c->level + 2, c);
	    w->root_state = w->states - 1;
	 }
	 /* locate_delta_images (wfalib.c:699-730, called at coder.c:876): the delta flags the
	    stream carries are derived from the structure, top down */
	 for (state = w->root_state; state >= w->basis_states; state--)
	    w->delta_state [state] = 0;
	 for (state = w->root_state; state >= w->basis_states; state--)
	    for (label = 0; label < MAXLABELS; label++)
	       if (w->tree [state][label] != RANGE
		   && (w->mv_type [state][label] != 0 || w->into [state][label][0] != NO_EDGE
		       || w->delta_state [state]))
		  w->delta_state [w->tree [state][label]] = 1;
	 memcpy (&out [coded], w, sizeof *w);

	 /* regenerate the frame: a reference of the frames to come (coder.c:642-651) */
	 {
	    int16_t *planes [3] = {cur, p->color ? cur + npix : NULL, p->color ? cur + 2 * npix : NULL};

	    fo_decode_image (w, p->color, (unsigned) p->width, (unsigned) p->height, planes);
	    if (type && p->color)
	       fo_restore_mc_colour (w, (unsigned) p->width, (unsigned) p->height, cur, past, fut);
	    else if (type)
	       fo_restore_mc (w, (unsigned) p->width, (unsigned) p->height, 0, cur, past, fut);
	    if (reconst)
	       memcpy (reconst + (size_t) coded * bands * npix, cur, bands * npix * sizeof (int16_t));
	 }
	 coded++;
	 remove_states (w->basis_states, w);
      }
      free (fut);
   }

cleanup:
   for (state = 0; state < MAXSTATES; state++)
   {
      free (c->images_of_state [state]);
      free (c->ip_images_state [state]);
      for (level = 0; level < MAXLEVEL; level++)
	 free (c->ip_states_state [state][level]);
   }
   for (level = 0; level < MAXLEVEL; level++)
   {
      free (c->mc_forward_norms [level]);
      free (c->mc_backward_norms [level]);
   }
   free (c->pixels);
   free (c);
   free (w);
   free (past);
   free (cur);
   free (chroma);
   return rc;
}

/* fill_norms_table (mwfa.c:544-602) for one block, exported for the parity test of the
   device's norms kernel */
void
fo_fill_norms_table (const int16_t *orig, const int16_t *past, unsigned width, unsigned height,
		     unsigned x0, unsigned y0, unsigned level, unsigned search_range, float *out)
{
   coder_t *c = calloc (1, sizeof (coder_t));

   c->planes [0]	       = orig;
   c->past		       = past;
   c->opt.width		       = (int) width;
   c->opt.height	       = (int) height;
   c->search_range	       = search_range;
   c->mc_forward_norms [level] = out;
   /* out-of-frame displacements are zeroed in both tables whatever the frame type */
   c->mc_backward_norms [level] = calloc ((size_t) 4 * search_range * search_range, sizeof (float));
   fill_norms_table (x0, y0, level, c);
   free (c->mc_backward_norms [level]);
   free (c);
}
