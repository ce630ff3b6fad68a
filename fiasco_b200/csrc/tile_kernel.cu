/*
 *  tile_kernel.cu -- the persistent FIASCO tile kernel for sm_100a.
 *
 *  One thread block encodes one independent tile / frame: it walks the whole bintree
 *  recursion of codec/subdivide.c on the device with an explicit DFS stack, and spreads
 *  the data-parallel parts over its threads:
 *
 *    - range x state products of an lc_max block (codec/ip.c:72 compute_ip_images_state):
 *	threads span the states, block pixels staged in shared memory;
 *    - state x state rows of a new state (codec/ip.c:184 compute_ip_states_state,
 *	codec/control.c:205 compute_images): threads span (level, partner state);
 *    - matching pursuit (codec/approx.c:317): threads span the domain pool; pass-1 bound,
 *	pass-2 quantised evaluation, an ordered warp-level resolution that reproduces the
 *	reference's index-ordered running minimum, and the Gram-Schmidt update
 *	(codec/approx.c:644) across the pool.
 *
 *  Many tiles run concurrently (grid = number of tiles).  All arithmetic that decides an
 *  index is plain fp32 with the reference's operation order (compiled with -fmad=false,
 *  IEEE division), double only inside log2, so the automaton is bit-identical to the
 *  reference CPU coder's.
 *
 *  Conventions inside this file: functions whose name starts with cta_ are executed by
 *  the whole block, begin after a barrier and end with a barrier.  t0_ functions are
 *  executed by thread 0 only, between barriers.
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stddef.h>
#include "tile_kernel.cuh"

#ifndef FB200_MINB
#define FB200_MINB 4		/* resident 128-thread blocks per SM the register budget is set for */
#endif

namespace {

/*****************************************************************************
				small helpers
*****************************************************************************/

__device__ __forceinline__ unsigned width_of_level (int l)  { return 1u << (l >> 1); }
__device__ __forceinline__ unsigned height_of_level (int l) { return 1u << ((l + 1) >> 1); }
__device__ __forceinline__ float fmin2 (float a, float b)   { return a > b ? b : a; }

/*
 *  The helpers below are deliberately NOT inlined: the kernel is one long serial chain whose
 *  speed is set by instruction fetch (ncu: "no instruction" is the top stall of the working
 *  warp), so one shared copy of log2 / rtob beats fifty inlined ones.
 */
__device__ __noinline__ double dev_log2d (float x)
{
   return log2 ((double) x);
}

/* -log2 (x) rounded to fp32 the way the reference does it: double log2, negate, narrow */
__device__ __forceinline__ float neg_log2f_via_double (float x)
{
   return (float) (-dev_log2d (x));
}

/* lib/rpf.c:59-111 (rtob); the reference shifts by more than 31 bits for tiny/huge
   inputs; on x86-64 that is a shift by (count mod 32), reproduced here explicitly */
__device__ __noinline__ int dev_rtob (float f, int mantissa_bits, float range)
{
   f = f / range;
   unsigned u	     = __float_as_uint (f);
   unsigned mantissa = u & 0x7fffffu;
   int	    exponent = (int) ((u >> 23) & 0xffu) - 126;
   int	    sign     = (int) (u >> 31);

   mantissa >>= 1;
   mantissa  |= 1u << 22;
   if (exponent > 0)
      mantissa <<= (exponent & 31);
   else
      mantissa >>= ((-exponent) & 31);
   mantissa >>= (23 - mantissa_bits - 1);
   mantissa  += 1;
   mantissa >>= 1;
   if (mantissa == 0)
      return -1;
   else if (mantissa >= (1u << mantissa_bits))
      return sign;
   else
      return (int) (((mantissa & ((1u << mantissa_bits) - 1)) << 1) | (unsigned) sign);
}

/* lib/rpf.c:113-169 (btor) */
__device__ __noinline__ float dev_btor (int binary, int mantissa_bits, float range)
{
   if (binary == -1)
      return 0.0f;
   int	    sign     = binary & 1;
   unsigned mantissa = (((unsigned) binary) & ((1u << (mantissa_bits + 1)) - 1)) >> 1;
   float    v;

   mantissa <<= (23 - mantissa_bits);
   if (mantissa == 0)
      v = sign ? -1.0f : 1.0f;
   else
   {
      int exponent = 0;

      while (!(mantissa & (1u << 22)))
      {
	 exponent--;
	 mantissa <<= 1;
      }
      mantissa <<= 1;
      v = __uint_as_float (((unsigned) sign << 31)
			   | ((unsigned) (exponent + 126) << 23)
			   | (mantissa & 0x7fffffu));
   }
   return v * range;
}

/* lib/misc.c:296-315 */
__device__ __forceinline__ unsigned dev_bits_bin_code (unsigned value, unsigned maxval)
{
   unsigned k = 31u - (unsigned) __clz ((int) (maxval + 1));
   unsigned r = (maxval + 1) - (1u << k);

   return value < maxval + 1 - 2 * r ? k : k + 1;
}

/*
 *  The tile's pointers are read from a table in shared memory, so the compiler cannot know
 *  that they point to global memory and would emit generic loads / stores -- which it must
 *  also keep in order with every shared-memory store.  GP () tells it.
 */
template <typename T>
__device__ __forceinline__ T *
as_global (T *p)
{
   __builtin_assume (__isGlobal (p));
   return p;
}
#define GP(p) as_global (p)

/*
 *  Thread-block cluster per stream (latency mode: fewer streams than SMs).  Rank 0 walks the
 *  recursion; the other blocks of the cluster wait at the cluster barrier for jobs: their share
 *  of the range x state products of a new block, of the state x state rows of a new state, and
 *  the pursuits of the label-0 descendants of a range, which see the same models and states as
 *  the range itself (codec/subdivide.c:188-237, 303-310).  Job descriptors, models and results
 *  travel through distributed shared memory (mapa + ordinary stores), ordered by
 *  barrier.cluster.arrive.release / wait.acquire, which also orders the global tables.
 */
#ifdef FB200_EMU
__device__ inline unsigned cl_rank (void) { return emu_cluster_rank (); }
__device__ inline void	   cl_sync (void) { emu_cluster_sync (); }
template <typename T> __device__ inline T *cl_map (T *p, unsigned rank) { return (T *) emu_map_shared_rank (p, rank); }
#else
__device__ __forceinline__ unsigned
cl_rank (void)
{
   unsigned r;
   asm volatile ("mov.u32 %0, %%cluster_ctarank;" : "=r" (r));
   return r;
}
__device__ __forceinline__ void
cl_sync (void)
{
   asm volatile ("barrier.cluster.arrive.release.aligned;\n\t"
		 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
/* the address of the same shared-memory object in the block of another rank (generic address
   of the shared::cluster window: ordinary loads and stores work on it) */
template <typename T>
__device__ __forceinline__ T *
cl_map (T *p, unsigned rank)
{
   unsigned long long out;
   asm volatile ("mapa.u64 %0, %1, %2;" : "=l" (out) : "l" ((unsigned long long) p), "r" (rank));
   return (T *) out;
}
#endif

/* only the shape that runs with an SM to itself can be clustered: the batch shapes (and the kernel
   of predicted frames) are compiled without any of it, so that their registers and their code
   size stay what they were */
template <int NT, bool MOTION>
__host__ __device__ constexpr bool
clustered_shape (void)
{
   (void) MOTION;
   return NT >= 512;
}

/* transitions of a state in registers */
struct TransReg
{
   int	 child [2];
   int	 into [2][FB_MAXEDGES];	/* FB_NO_EDGE from the first unused slot on */
   float w [2][FB_MAXEDGES];
};

__device__ __forceinline__ void
load_trans (const Trans *p, TransReg &r)
{
   const uint4 *q = (const uint4 *) p;
   const uint4	a = q [0], b = q [1], c = q [2], d = q [3];	/* one 64-byte line */
   const unsigned sh [12] = {a.x, a.y, a.z, a.w, b.x, b.y};
   short	  v [12];

#pragma unroll
   for (int i = 0; i < 6; i++)
   {
      v [2 * i]	    = (short) (sh [i] & 0xffffu);
      v [2 * i + 1] = (short) (sh [i] >> 16);
   }
   r.child [0] = v [0];
   r.child [1] = v [1];
   bool end0 = false, end1 = false;
#pragma unroll
   for (int e = 0; e < FB_MAXEDGES; e++)
   {
      end0 = end0 || v [2 + e] == FB_NO_EDGE;
      end1 = end1 || v [7 + e] == FB_NO_EDGE;
      r.into [0][e] = end0 ? FB_NO_EDGE : (int) v [2 + e];
      r.into [1][e] = end1 ? FB_NO_EDGE : (int) v [7 + e];
   }
   r.w [0][0] = __uint_as_float (b.z);
   r.w [0][1] = __uint_as_float (b.w);
   r.w [0][2] = __uint_as_float (c.x);
   r.w [0][3] = __uint_as_float (c.y);
   r.w [0][4] = __uint_as_float (c.z);
   r.w [1][0] = __uint_as_float (c.w);
   r.w [1][1] = __uint_as_float (d.x);
   r.w [1][2] = __uint_as_float (d.y);
   r.w [1][3] = __uint_as_float (d.z);
   r.w [1][4] = __uint_as_float (d.w);
}

/* thread 0: pack the transitions of state s (already in the automaton arrays) */
__device__ void
t0_store_trans (const TileWs &W, unsigned s)
{
   Trans t;

   for (int label = 0; label < 2; label++)
   {
      const short *in = GP (W.into) + (size_t) (2 * s + label) * 6;
      const float *wt = GP (W.weight) + (size_t) (2 * s + label) * 6;
      bool	   end = false;

      t.child [label] = GP (W.tree) [2 * s + label];
      for (int e = 0; e < FB_MAXEDGES; e++)
      {
	 end = end || in [e] == FB_NO_EDGE;
	 t.into [label][e] = end ? (short) FB_NO_EDGE : in [e];
	 t.w [label][e]	   = end ? 0.0f : wt [e];
      }
   }
   GP (W.trans) [s] = t;
}

/*****************************************************************************
			     shared memory layout
*****************************************************************************/

struct RangeRes			/* the range_t fields the still-image path uses */
{
   float	  weight [FB_MAXEDGES + 1];
   float	  err, tree_bits, matrix_bits, weights_bits;
   short	  into [FB_MAXEDGES + 1];
   short	  tree;
   unsigned short x, y;
};

struct Frame			/* one activation record of subdivide() */
{
   float    max_costs, lincomb_costs, subdivide_costs;
   unsigned x, y, image, address;
   unsigned gaddr;		/* range->global_address: address in the whole picture (progress meter) */
   int	    level, y_state, label;
   int	    spec_k;		/* position in the current spine of speculated pursuits, or -1 */
   short    lc_code [FB_MAXEDGES];	/* quantiser codes of lrange's weights and the y-state its pursuit */
   short    lc_ystate;		/* saw: the models take the range when it wins (ST_DECIDE) */
   int	    new_y_state [2];
   unsigned states_snap;
   float    r_err, r_tree_bits, r_matrix_bits, r_weights_bits; /* rrange sums */
   RangeRes lrange;
   RangeRes child [2];
};

/* what a range of a predicted frame carries beside RangeRes (range_t, codec/cwfa.h:46-75) */
struct RangeX
{
   float       mv_tree_bits, mv_coord_bits;
   signed char mv_type, mv_fx, mv_fy, mv_bx, mv_by, prediction;
};

/* activation record of subdivide() in a predicted frame, beside Frame */
struct FrameX
{
   RangeX   lrange, child [2], prange_x;
   RangeRes prange;		/* result of the nested pass over the prediction error */
   float    r_mvt, r_mvc;	/* rrange sums of the motion bits */
   int	    try_mc;		/* subdivide.c:141-147 */
   int	    try_nd;		/* subdivide.c:149-151: an intra frame coded with `--prediction' */
   float    ndw;		/* weight of the DC component that predicts the range (prediction.c:385) */
   int	    saved_off;		/* hole_off before the prediction attempt */
   int	    off_snap;		/* hole_off when the range was entered (beside Frame.states_snap) */
   int	    delta, prediction;	/* arguments of subdivide () */
   int	    pred_done;		/* the prediction alternative has been tried (and lost) */
   unsigned rec_states;		/* states after the first two alternatives */
   unsigned last_state;
   float    max_pred, pcosts, mvc, mvt;
   int	    mx, my, bx, by, mctype;
};

struct MpRes			/* mp_t, codec/approx.c:41-51 */
{
   short indices [FB_MAXEDGES + 1];
   short into [FB_MAXEDGES + 1];
   float weight [FB_MAXEDGES];
   short code [FB_MAXEDGES];	/* rtob (weight) with the rpf of the domain's kind */
   float matrix_bits, weights_bits, err, costs;
};

struct MpWork
{
   float  B [FB_MAXEDGES];	/* ip_image_ortho_vector */
   float  N [FB_MAXEDGES];	/* norm_ortho_vector */
   float  fB [FB_MAXEDGES];	/* B[k] / N[k] */
   float  mbase [FB_MAXEDGES + 2]; /* -log2 (count[k] / total) */
   float  d0b [2];		/* DC-column bits: [0] unused, [1] used */
   double *l2_dc;		/* [aac_dc_size] log2 (count / total) per DC code */
   double *l2_lv;		/* [aac_lvl_size] same for the context of the current level */
   float  additional_bits, price, norm, min_costs;
   float  wb_dc, wb_nd;		/* pass-1 weights bits of a DC / other candidate */
   int	  nc;			/* chosen vectors with non-zero weight */
   short  cvec [FB_MAXEDGES + 1];
   short  csorted [FB_MAXEDGES + 1];
   int	  ncs;			/* of which not the y-state */
   int	  n;			/* current step */
   int	  D, pool_n, ydom, y_state, level, li, size;
   int	  index;		/* winner of the current step or -1 */
   int	  wave_pos, wave_done;	/* lazy pass-2 waves: next start index, finished flag */
   float  best_f [FB_MAXEDGES];
   short  best_c [FB_MAXEDGES];
   short  half_lv, half_dc;	/* rtob (0.5) of the two quantisers */
   float  best_mbits, best_wbits, best_err, best_costs;
   unsigned st_steps, st_pass2;	/* work counters of this pursuit (added to the tile's when it is used) */
   int	  win;			/* wide waves: rank of the winning candidate or -1 */
};

/* jobs of the helper blocks of a cluster */
#define FB_SPINE_MAX 8
#define FB_WIDE_PAYLOAD 13	/* 5 weights, 5 codes, matrix bits, weights bits, error */
enum { CJ_EXIT, CJ_SPINE, CJ_TINIT, CJ_APPEND, CJ_TERR };

struct SpineNode		/* one pursuit of a spine: the range (level, image, address) */
{
   int	    level;
   unsigned image, address;
   float    tree_bits, norm;
   float    mv_tree_bits;	/* predicted frames: 1 if motion compensation is allowed for the range */
};

struct ClJob
{
   int	     type, n;		/* CJ_*; SPINE: number of nodes */
   unsigned  states;		/* wfa->states */
   unsigned  pool_n;		/* SPINE: entries of the pool list */
   int	     x, y, band;	/* TINIT: the block */
   unsigned  s;			/* APPEND: the new state */
   float     price;
   int	     nested;		/* predicted frames: the pass over a prediction error (delta models, its pixels) */
   int	     tswap;		/* predicted frames: the product tables T / T2 are swapped */
   int	     level;		/* TERR: level of the prediction-error block */
   SpineNode node [FB_SPINE_MAX];
};

struct SpecRes			/* result of a speculated pursuit */
{
   MpRes    mp;
   unsigned steps, pass2, D;
};

struct ShHdr
{
   int	    state [2];
   int	    depthv [2];
   int	    status;
   int	    band;
   float    ret_costs;
   float    price;		/* price of the current band */
   unsigned y_states;		/* states at the end of the luminance band */
   int	    lc_min;		/* current lc_min_level (raised for the chroma bands, coder.c:785-797) */
   unsigned states;		/* wfa->states */
   unsigned tree_counts [FB200_MAXLEVEL];
   unsigned tree_total [FB200_MAXLEVEL];
   int	    trace_len;
   unsigned long long mp_calls, mp_steps, pass2, blocks, ip_bytes, mp_bytes, ss_bytes;
   long long cyc_T, cyc_mp, cyc_append, cyc_start;
   long long lap [16], lap_last;	/* thread-0 lap timer per sub-phase */
   /* sources of the state being appended (cta_state_products) */
   short    ap_dom [2][FB_MAXEDGES + 1];
   float    ap_w [2][FB_MAXEDGES + 1];
   signed char ap_row [2][FB_MAXEDGES + 1];
   unsigned char ap_cnt [2], ap_child [2];
   short    ap_src [2 * (FB_MAXEDGES + 1)];
   int	    ap_nsrc;
   MpRes    mp, tmp;
   MpWork   w;
   RangeRes root;
   Frame   *frames;		/* [level - lc_min + 2] */
   /* predicted frames */
   FrameX  *fx;			/* [n_frames] or NULL */
   RangeX   root_x;
   int	    nest_base;		/* depth of the root record of the nested pass, or -1 */
   int	    top;		/* level of node 0 of the product tree in use */
   int	    best_i;		/* find_best_mv: winning displacement index */
   float    best_c;		/* and its costs */
   float    fcosts, bcosts;	/* find_B_frame_mc */
   int	    fi, bi;
   long long isum;		/* integer sum of squares of the interpolated prediction error */
   /* the percent meter of subdivide() (subdivide.c:105-108,323-337): last value, values shown */
   unsigned percent, progress [4];
   /* cluster per stream */
   ClJob    job;		/* rank 0: the job being posted; helpers: the job received */
   SpecRes  spec [FB_SPINE_MAX];	/* rank 0: pursuits of the current spine (entry 0 unused) */
   int	    spec_len;		/* nodes of the current spine */
   float    spec_rbits [FB_SPINE_MAX];	/* tree bits of "subdivided" at the levels of its nodes */
   int	    pool_lo;		/* lowest pool-list entry written since the list was last posted */
   int	    tswap;		/* predicted frames: W.T and W.T2 are swapped (nested pass) */
   /* bulk copies (cp.async.bulk, the TMA unit) of table rows into shared memory complete on this */
   unsigned long long mbar;
   unsigned mbar_phase;
   /* colour frames of a sequence: what the reference's wfa->y_column array holds (yc_ref below) */
   int	    hole_off;		/* device state number - this = the state's number in the reference */
   unsigned ref_max;		/* highest number of states the reference's luminance band reaches */
};

static_assert (offsetof (ShHdr, tree_total) == offsetof (ShHdr, tree_counts) + FB200_MAXLEVEL * sizeof (unsigned),
	       "the tree model is copied as one array of 2 * MAXLEVEL counters");

struct Sh			/* pointers into dynamic shared memory */
{
   ShHdr   *h;
   float   *num, *den, *G;	/* [dcap], [dcap], [5][dcap] */
   unsigned char *used;		/* [dcap] */
   short   *pool;		/* [s_cap] domain index -> state */
   float   *pixels;		/* [2^lc_max] */
   int	   *norm_i;		/* [tn] integer sum of squares per node */
   float   *bnd;		/* [dcap32] pass-1 bound per domain (INF: not usable) */
   unsigned *cmask;		/* [dcap32 / 32] candidate bit masks of the current wave */
   short   *blob;		/* [blob_len] current probability models */
   short   *snaps;		/* [ndepth][2][blob_len] model snapshots of the DFS, or NULL */
   double  *l2;			/* [aac_dc_size + aac_lvl_size] */
   float   *qt_dc, *qt_lv;	/* [aac_dc_size], [aac_lvl_size]: btor (code), lib/rpf.c:113 */
   unsigned *tsnap;		/* [ndepth][2 * MAXLEVEL] tree-model snapshots of the DFS */
   Frame   *frames;
   FrameX  *fx;			/* predicted frames only */
   int	    dcap;
   int	    scratch_len;	/* floats from num to the end of bnd / G: row scratch of append_state */
   float   *wide;		/* [2 + FB_WIDE_PAYLOAD][NT] candidates of a wide wave (NT >= 512 only) */
};

__host__ __device__ inline size_t align16 (size_t x) { return (x + 15) & ~(size_t) 15; }

__host__ __device__ inline size_t
smem_layout (const DevParams &p, int nt, size_t *off /* [20] */)
{
   size_t o    = 0;
   size_t dcap = (size_t) p.s_cap + 1;

   off [0] = o; o += align16 (sizeof (ShHdr));
   off [1] = o; o += align16 (dcap * 4);			/* num */
   off [2] = o; o += align16 (dcap * 4);			/* den */
   off [8] = o; o += align16 (((dcap + 31) / 32 * 32) * 4);	/* bnd */
   off [3] = o; o += (p.big & 1) ? 0 : align16 (dcap * 4 * (p.max_elements > 1 ? p.max_elements - 1 : 1)); /* G */
   off [4] = o; o += align16 (dcap);				/* used */
   off [5] = o; o += align16 ((size_t) p.s_cap * 2);		/* pool */
   off [6] = o; o += align16 (((size_t) 1 << p.lc_max) * 4);	/* pixels */
   off [7] = o; o += align16 ((size_t) p.tn * 4);		/* norm_i */
   off [9] = o; o += align16 (((dcap + 31) / 32) * 4);		/* cmask */
   off [10] = o; o += align16 ((size_t) p.blob_len * 2);	/* blob */
   {
      /* model snapshots of the DFS stay on chip when they are small (default models:
	 376 B each), else they live in the tile's global workspace */
      const size_t need = (size_t) p.n_frames * (p.motion ? 3 : 2) * p.blob_len * 2;
      const bool on_chip = need <= 24 * 1024 && !(p.big & 2);
      off [11] = on_chip ? o : (size_t) -1;
      if (on_chip)
	 o += align16 (need);
   }
   off [12] = o; o += align16 ((size_t) (p.aac_dc_size + p.aac_lvl_size) * 8);	/* log2 tables */
   off [13] = o; o += align16 ((size_t) p.n_frames * sizeof (Frame));	/* DFS frames */
   off [14] = o; o += align16 ((size_t) (p.aac_dc_size + p.aac_lvl_size) * 4);	/* quantiser tables */
   off [15] = o; o += align16 ((size_t) p.n_frames * (p.motion ? 2 : 1) * 2 * FB200_MAXLEVEL * 4); /* tsnap */
   off [16] = o; o += p.motion ? align16 ((size_t) p.n_frames * sizeof (FrameX)) : 0;
   off [17] = o; o += align16 (sizeof (TileWs));	/* the tile's pointer table */
   /* the shape with an SM to itself evaluates up to nt candidates of a pursuit step at once: key,
      costs and the 13 words of a candidate's result, one column per thread (cta_mp_find_wide) */
   off [18] = o; o += nt >= 512 ? align16 ((size_t) nt * 4 * (2 + FB_WIDE_PAYLOAD)) : 0;
   off [19] = o;
   return o;
}

__device__ __forceinline__ Sh
carve (unsigned char *base, const DevParams &p, int nt, float *gglob)
{
   /* the offsets come ready-made from the host (launch_nt): what the compiler rematerialises
      when it is short of registers is then one constant-bank load, not the layout arithmetic */
   const unsigned *off = p.sm_off;
   Sh		   s;

   s.h	    = (ShHdr *) (base + off [0]);
   s.num    = (float *) (base + off [1]);
   s.den    = (float *) (base + off [2]);
   s.G	    = (p.big & 1) ? gglob : (float *) (base + off [3]);
   s.used   = (unsigned char *) (base + off [4]);
   s.pool   = (short *) (base + off [5]);
   s.pixels = (float *) (base + off [6]);
   s.norm_i = (int *) (base + off [7]);
   s.bnd    = (float *) (base + off [8]);
   s.cmask  = (unsigned *) (base + off [9]);
   s.blob   = (short *) (base + off [10]);
   s.snaps  = off [11] == 0xffffffffu ? (short *) 0 : (short *) (base + off [11]);
   s.dcap   = p.s_cap + 1;
   s.scratch_len = (int) ((off [4] - off [1]) / 4);
   s.l2	    = (double *) (base + off [12]);
   s.frames = (Frame *) (base + off [13]);
   s.qt_dc  = (float *) (base + off [14]);
   s.qt_lv  = s.qt_dc + p.aac_dc_size;
   s.tsnap  = (unsigned *) (base + off [15]);
   s.fx	    = p.motion ? (FrameX *) (base + off [16]) : (FrameX *) 0;
   s.wide   = (float *) (base + off [18]);
   return s;
}

/* accessors of the model blob */
#define BLOB_U16(sh, i) (((unsigned short *) (sh).blob) [i])
#define BLOB_S16(sh, i) ((sh).blob [i])

/* thread-0 lap timer: time since the previous LAP goes to bucket i */
#ifndef FB200_LAPS	/* diagnostics build only: the timers cost 2-3 % */
#define LAP(h, i) do { } while (0)
#else
#define LAP(h, i) do { if (threadIdx.x == 0) { const long long now_ = clock64 (); (h)->lap [i] += now_ - (h)->lap_last; (h)->lap_last = now_; } } while (0)
#endif
#ifdef FB200_X_WAVELAPS	/* experiment: the three product buckets time the parts of a wide wave instead */
#define LAP_W(h, i) LAP (h, i)
#define LAP_T(h, i) LAP (h, LAP_CTRL)
#else
#define LAP_W(h, i) do { } while (0)
#define LAP_T(h, i) LAP (h, i)
#endif
enum { LAP_CTRL, LAP_PIX, LAP_DOTS, LAP_UPSWEEP, LAP_ENTER, LAP_MP_PRO, LAP_MP_P1, LAP_MP_WAVES,
       LAP_MP_COMMIT, LAP_MP_ORTHO, LAP_AR_EPI, LAP_AP_IMG, LAP_AP_DIRECT, LAP_AP_STAGED, LAP_DECIDE, LAP_CLUSTER, LAP_N };

enum { ST_ENTER, ST_CHILD, ST_AFTER_CHILD, ST_DECIDE, ST_RETURN, ST_DONE, ST_ABORT };

/*****************************************************************************
		    probability models (thread-0 scalar code)
*****************************************************************************/

/* codec/bintree.c:55-68 */
/* (the kernel of predicted frames keeps the prediction tree model, c->p_tree, in the upper halves
   of the same words: one snapshot holds both; a level's counts stay below 2 * FB200_MAXSTATES) */
__device__ float t0_tree_bits (const ShHdr *h, int child, int level)
{
   float prob = (h->tree_counts [level] & 0xffffu) / (float) (h->tree_total [level] & 0xffffu);

   return child ? neg_log2f_via_double (prob) : neg_log2f_via_double (1 - prob);
}

/* the same for the prediction tree model (subdivide.c:259-260, prediction.c:391) */
__device__ float t0_ptree_bits (const ShHdr *h, int child, int level)
{
   float prob = (h->tree_counts [level] >> 16) / (float) (h->tree_total [level] >> 16);

   return child ? neg_log2f_via_double (prob) : neg_log2f_via_double (1 - prob);
}

/* qac_bits (domain-pool.c:367-402) as rle_bits calls it on the 1-entry DC model */
__device__ float t0_d0_bits (const Sh &sh, int dc_used, int y_state,
			     const float *m0, const float *m1)
{
   float bits = 0;

   if (BLOB_U16 (sh, MB_D0N) > 0 && 0 != y_state)
      bits += m0 [BLOB_S16 (sh, MB_D0INDEX)];
   if (y_state >= 0)
      bits += m0 [BLOB_U16 (sh, MB_D0YINDEX)];
   if (dc_used)
   {
      if (0 == y_state)
      {
	 bits -= m0 [BLOB_U16 (sh, MB_D0YINDEX)];
	 bits += m1 [BLOB_U16 (sh, MB_D0YINDEX)];
      }
      else
      {
	 bits -= m0 [BLOB_S16 (sh, MB_D0INDEX)];
	 bits += m1 [BLOB_S16 (sh, MB_D0INDEX)];
      }
   }
   return bits;
}

/*
 *  rle_bits (domain-pool.c:737-793) for a used-domain list that is already sorted and
 *  already stripped of the y-state domain.  mbase / d0b are per-call tables.  M bounds
 *  the list length at compile time.
 */
template <int M>
__device__ __forceinline__ float
dev_rle_bits (const float *mbase, const float *d0b, const short *sorted, int n,
	      unsigned pool_n)
{
   float    bits = mbase [n];
   unsigned last = 1;

   bits += (n && sorted [0] == 0) ? d0b [1] : d0b [0];
#pragma unroll
   for (int e = 0; e < M; e++)
      if (e < n)
      {
	 int into = sorted [e];

	 if (into && (pool_n - 1 - last))
	 {
	    bits += (float) dev_bits_bin_code ((unsigned) into - last, pool_n - 1 - last);
	    last  = (unsigned) into + 1;
	 }
      }
   return bits;
}

/* insert d into the ascending list srt[0..ns) of capacity M */
template <int M>
__device__ __forceinline__ void
sorted_insert (short (&srt) [M], int &ns, int d)
{
   int p = ns;

#pragma unroll
   for (int e = M - 1; e > 0; e--)
      if (e <= p && srt [e - 1] > d)
      {
	 srt [e] = srt [e - 1];
	 p	 = e - 1;
      }
#pragma unroll
   for (int e = 0; e < M; e++)
      if (e == p)
	 srt [e] = (short) d;
   ns++;
}

__constant__ float c_matrix_0 [1024];	/* domain-pool.c:970-999 */
__constant__ float c_matrix_1 [1024];

/*****************************************************************************
		range x state products of a block  (codec/ip.c:72-154)
*****************************************************************************/

/*
 *  Upsweep of U consecutive nodes of one level for state s (ip.c:98-151): the gathers of all
 *  U nodes are issued before the first addition, so their L2 round trips overlap; each
 *  entry is still accumulated in the reference's order.
 */
template <int U>
__device__ __forceinline__ void
upsweep_nodes (float *T, unsigned scap, unsigned node, unsigned s, const TransReg &tr,
	       const int (&idx) [2][FB_MAXEDGES + 1])
{
   float v [U][2][FB_MAXEDGES + 1];

#pragma unroll
   for (int u = 0; u < U; u++)
#pragma unroll
      for (int label = 0; label < 2; label++)
      {
	 const float *src = T + (size_t) (2 * (node + u) + 1 + label) * scap;
#pragma unroll
	 for (int e = 0; e < FB_MAXEDGES + 1; e++)
	    v [u][label][e] = src [idx [label][e]];
      }
#pragma unroll
   for (int u = 0; u < U; u++)
   {
      float acc = 0;
#pragma unroll
      for (int label = 0; label < 2; label++)
      {
	 if (tr.child [label] != FB_RANGE)
	    acc += v [u][label][0];
#pragma unroll
	 for (int e = 0; e < FB_MAXEDGES; e++)
	    if (tr.into [label][e] != FB_NO_EDGE)
	       acc += v [u][label][e + 1] * tr.w [label][e];
      }
      T [(size_t) (node + u) * scap + s] = acc;
   }
}

/*
 *  Block-wide asynchronous copy of nfloats (a multiple of 4) floats from global to shared
 *  memory, both 16-byte aligned: cp.async moves 16 bytes per instruction past the register
 *  file and leaves the thread free to issue the next one; the caller commits and waits.
 */
template <int NT>
__device__ __forceinline__ void
cta_copy_f32_async (float *dst_smem, const float *src, unsigned nfloats)
{
#ifdef FB200_EMU		/* tests/emu: the copy is synchronous */
   for (unsigned i = threadIdx.x; i < (nfloats >> 2); i += NT)
      ((uint4 *) dst_smem) [i] = ((const uint4 *) src) [i];
#else
   const unsigned saddr = (unsigned) __cvta_generic_to_shared (dst_smem);

   for (unsigned i = threadIdx.x; i < (nfloats >> 2); i += NT)
      asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r" (saddr + i * 16), "l" (src + i * 4) : "memory");
#endif
}

/* one float from global to shared memory, asynchronously (a gather that costs no register
   and does not make the thread wait: all gathers of a thread are in flight together) */
__device__ __forceinline__ void
gather_f32_async (float *dst_smem, const float *src)
{
#ifdef FB200_EMU
   *dst_smem = *src;
#else
   asm volatile ("cp.async.ca.shared.global [%0], [%1], 4;"
		 :: "r" ((unsigned) __cvta_generic_to_shared (dst_smem)), "l" (src) : "memory");
#endif
}

__device__ __forceinline__ void
async_wait_all (void)
{
#ifndef FB200_EMU
   asm volatile ("cp.async.commit_group;" ::: "memory");
   asm volatile ("cp.async.wait_group 0;" ::: "memory");
#endif
}

/*
 *  Bulk copies global -> shared by the TMA unit (cp.async.bulk: one instruction per row instead of
 *  one 16-byte cp.async per thread and 16 bytes), completion through an mbarrier in shared memory.
 *  Rows are multiples of 16 bytes, 16-byte aligned at both ends.
 */
__device__ __forceinline__ void
mbar_init (unsigned long long *mbar)
{
#ifndef FB200_EMU
   asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r" ((unsigned) __cvta_generic_to_shared (mbar)) : "memory");
   asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
   *mbar = 0;
#endif
}

/* one thread: announce the bytes of the copies that follow */
__device__ __forceinline__ void
mbar_expect (unsigned long long *mbar, unsigned bytes)
{
#ifndef FB200_EMU
   asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		 :: "r" ((unsigned) __cvta_generic_to_shared (mbar)), "r" (bytes) : "memory");
#endif
}

__device__ __forceinline__ void
bulk_copy_row (float *dst_smem, const float *src, unsigned bytes, unsigned long long *mbar)
{
#ifdef FB200_EMU
   memcpy (dst_smem, src, bytes);
#else
   asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		 :: "r" ((unsigned) __cvta_generic_to_shared (dst_smem)), "l" (src), "r" (bytes),
		    "r" ((unsigned) __cvta_generic_to_shared (mbar)) : "memory");
#endif
}

/* every thread: wait until the copies of the current phase have landed */
__device__ __forceinline__ void
mbar_wait (unsigned long long *mbar, unsigned phase)
{
#ifndef FB200_EMU
   unsigned done;

   do
      asm volatile ("{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r" (done) : "r" ((unsigned) __cvta_generic_to_shared (mbar)), "r" (phase) : "memory");
   while (!done);
#endif
}

/*
 *  Fill T[node][s] for the nodes of the subtree rooted at (node_root, level_root) of the
 *  current lc_max block and every state s >= from whose image is needed.
 *  Levels <= il are direct dot products with the state images (ip.c:268-295), higher
 *  levels the weighted sums over the state's transitions (ip.c:98-151), each entry
 *  accumulated in the reference's order: label 0 child, label 0 edges, label 1 ...
 */
template <int NT, bool CLU = false>
__device__ void
cta_compute_T (const DevParams &P, const TileWs &W, const Sh &sh, unsigned from,
	       unsigned node_root, int level_root, int top, unsigned g0 = 0, unsigned gnt = NT)
{
   /* g0, gnt: the block's first thread and the number of threads when the blocks of a cluster
      share the work (gnt > NT): the loops are strided over all of them and the levels are
      separated by the cluster barrier */
   const unsigned tid	= g0 + threadIdx.x;
   const bool	 cl	= CLU && gnt > (unsigned) NT;
   const unsigned S	= sh.h->states;
   const unsigned scap	= (unsigned) P.s_cap;

   if (from >= S)
      return;			/* uniform: nothing to do, no barrier needed */
   LAP (sh.h, LAP_CTRL);

   /* direct levels */
   for (int l = P.lmin; l <= P.il && l <= level_root; l++)
   {
      const unsigned nn	   = 1u << (level_root - l);
      const unsigned node0 = ((node_root + 1) << (level_root - l)) - 1;
      const unsigned adr0  = node0 - ((1u << (top - l)) - 1);
      const unsigned len   = 1u << l;
      const unsigned ns	   = S - from;

      if (ns >= 32)
      {
	 /* a warp covers 32 consecutive states and walks the nodes together: pixel
	    reads are shared-memory broadcasts, product writes are coalesced.  With the
	    threads of a whole cluster at hand the nodes of a state are split up as well, so
	    that every thread has an item */
	 const unsigned ns32 = (ns + 31) & ~31u;
	 unsigned	split = 1;

	 if (cl)
	    while (split < nn && ns32 * split < gnt)
	       split <<= 1;
	 const unsigned per = nn / split;

	 for (unsigned item = tid; item < ns32 * split; item += gnt)
	 {
	    const unsigned s  = from + item % ns32;
	    const unsigned k0 = (item / ns32) * per;

	    if (s >= S || !GP (W.domain_type) [s])
	       continue;
	    const float *im = GP (W.img) + (size_t) s * FB_IMG_STRIDE + (len - 1);
	    if (l == 5)
	    {
	       float r [32];
	       const float4 *im4 = (const float4 *) (GP (W.img) + (size_t) s * FB_IMG_STRIDE + 32);
	       /* level-5 part starts at float 31; rows are 256 B aligned, so read the
		  aligned float4s 7..15 and shift by one */
	       float prev = GP (W.img) [(size_t) s * FB_IMG_STRIDE + 31];
#pragma unroll
	       for (int q = 0; q < 8; q++)
	       {
		  float4 v = im4 [q];
		  r [4 * q]	= prev;
		  r [4 * q + 1] = v.x;
		  r [4 * q + 2] = v.y;
		  r [4 * q + 3] = v.z;
		  prev		= v.w;
	       }
	       for (unsigned k = k0; k < k0 + per; k++)
	       {
		  const float4 *px = (const float4 *) (sh.pixels + (size_t) (adr0 + k) * 32);
		  float		ip = 0;
#pragma unroll
		  for (int q = 0; q < 8; q++)
		  {
		     const float4 p4 = px [q];
		     ip += p4.x * r [4 * q];
		     ip += p4.y * r [4 * q + 1];
		     ip += p4.z * r [4 * q + 2];
		     ip += p4.w * r [4 * q + 3];
		  }
		  GP (W.T) [(size_t) (node0 + k) * scap + s] = ip;
	       }
	    }
	    else
	    {
	       for (unsigned k = k0; k < k0 + per; k++)
	       {
		  const float *px = sh.pixels + (size_t) (adr0 + k) * len;
		  float	       ip = 0;
		  for (unsigned i = 0; i < len; i++)
		     ip += px [i] * im [i];
		  GP (W.T) [(size_t) (node0 + k) * scap + s] = ip;
	       }
	    }
	 }
      }
      else
      {
	 /* few new states: spread (state, node) pairs over the threads */
	 for (unsigned item = tid; item < ns * nn; item += gnt)
	 {
	    const unsigned s = from + item / nn;
	    const unsigned k = item % nn;

	    if (!GP (W.domain_type) [s])
	       continue;
	    const float *im = GP (W.img) + (size_t) s * FB_IMG_STRIDE + (len - 1);
	    const float *px = sh.pixels + (size_t) (adr0 + k) * len;
	    float	 ip = 0;
	    for (unsigned i = 0; i < len; i++)
	       ip += px [i] * im [i];
	    GP (W.T) [(size_t) (node0 + k) * scap + s] = ip;
	 }
      }
   }
   if (cl)
      cl_sync ();
   else
      __syncthreads ();
   LAP_T (sh.h, LAP_DOTS);

   /* upsweep */
   for (int l = (P.il + 1 > P.lmin ? P.il + 1 : P.lmin); l <= level_root; l++)
   {
      const unsigned nn	   = 1u << (level_root - l);
      const unsigned node0 = ((node_root + 1) << (level_root - l)) - 1;
      const unsigned ns	   = S - from;

      if (ns >= gnt / 2)
      {
	 /* many states: a thread keeps one state's transitions in registers and walks
	    the nodes of this level; the gathers hit the two child rows of a node */
	 for (unsigned s = from + tid; s < S; s += gnt)
	 {
	    if (!GP (W.domain_type) [s])
	       continue;
	    TransReg tr;

	    load_trans (GP (W.trans) + s, tr);
	    /* gather indices with the unused slots pointing at entry 0 (always valid): the
	       loads are unconditional and independent, so several nodes' worth of L2 round
	       trips overlap; only the additions are predicated (the value stays exact) */
	    int idx [2][FB_MAXEDGES + 1];
#pragma unroll
	    for (int label = 0; label < 2; label++)
	    {
	       idx [label][0] = tr.child [label] != FB_RANGE ? tr.child [label] : 0;
#pragma unroll
	       for (int e = 0; e < FB_MAXEDGES; e++)
		  idx [label][e + 1] = tr.into [label][e] != FB_NO_EDGE ? tr.into [label][e] : 0;
	    }
	    if (nn >= 2)
#pragma unroll 1
	       for (unsigned k = 0; k < nn; k += 2)
		  upsweep_nodes<2> (GP (W.T), scap, node0 + k, s, tr, idx);
	    else
#pragma unroll 1
	       for (unsigned k = 0; k < nn; k++)
		  upsweep_nodes<1> (GP (W.T), scap, node0 + k, s, tr, idx);
	 }
      }
      else
      {
	 for (unsigned item = tid; item < ns * nn; item += gnt)
	 {
	    const unsigned s	= from + item % ns;
	    const unsigned node = node0 + item / ns;
	    TransReg	   tr;

	    if (CLU)
	    {
	       /* (the state's type and its transitions in flight together: one round trip less) */
	       const unsigned dt = GP (W.domain_type) [s];

	       load_trans (GP (W.trans) + s, tr);
	       if (!dt)
		  continue;
	    }
	    else
	    {
	       if (!GP (W.domain_type) [s])
		  continue;
	       load_trans (GP (W.trans) + s, tr);
	    }
	    if (CLU)
	    {
	       /* (unconditional, independent gathers as above: one L2 round trip per item; the
		  batch shapes, whose frames compete for L2, only load what exists) */
	       int idx [2][FB_MAXEDGES + 1];
#pragma unroll
	       for (int label = 0; label < 2; label++)
	       {
		  idx [label][0] = tr.child [label] != FB_RANGE ? tr.child [label] : 0;
#pragma unroll
		  for (int e = 0; e < FB_MAXEDGES; e++)
		     idx [label][e + 1] = tr.into [label][e] != FB_NO_EDGE ? tr.into [label][e] : 0;
	       }
	       upsweep_nodes<1> (GP (W.T), scap, node, s, tr, idx);
	    }
	    else
	    {
	       float acc = 0;
#pragma unroll
	       for (int label = 0; label < 2; label++)
	       {
		  const float *src = GP (W.T) + (size_t) (2 * node + 1 + label) * scap;

		  if (tr.child [label] != FB_RANGE)
		     acc += src [tr.child [label]];
#pragma unroll
		  for (int e = 0; e < FB_MAXEDGES; e++)
		     if (tr.into [label][e] != FB_NO_EDGE)
			acc += src [tr.into [label][e]] * tr.w [label][e];
	       }
	       GP (W.T) [(size_t) node * scap + s] = acc;
	    }
	 }
      }
      if (cl)
	 cl_sync ();
      else
	 __syncthreads ();
   }
   LAP_T (sh.h, LAP_UPSWEEP);
}

/* codec/subdivide.c:612-644 (init_range) + :504-541 (cut_to_bintree) */
template <int NT, bool CLU = false>
__device__ void
cta_init_range (const DevParams &P, const TileWs &W, const Sh &sh, unsigned x0,
		unsigned y0, int band, unsigned g0 = 0, unsigned gnt = NT)
{
   const int	  tid  = threadIdx.x;
   const unsigned size = 1u << P.lc_max;
   const int16_t *src  = GP (W.pix) + (size_t) band * P.width * P.height;

   for (unsigned i = tid; i < size; i += NT)
   {
      /* bintree order: y in the even address bits, x in the odd ones */
      unsigned yy = 0, xx = 0;
      for (int b = 0; b < 11; b++)
      {
	 yy |= ((i >> (2 * b)) & 1u) << b;
	 xx |= ((i >> (2 * b + 1)) & 1u) << b;
      }
      const unsigned px = x0 + xx, py = y0 + yy;
      float	     v	= 0;

      if (py < (unsigned) P.height && px < (unsigned) P.width)
	 v = (float) ((int) src [(size_t) py * P.width + px] / 16);
      sh.pixels [i] = v;
   }
   __syncthreads ();

   /* integer sums of squares per node (pixels are integers here); bottom level first */
   {
      const unsigned nleaf = 1u << (P.lc_max - P.lmin);
      const unsigned len   = 1u << P.lmin;

      for (unsigned k = tid; k < nleaf; k += NT)
      {
	 int acc = 0;
	 for (unsigned i = 0; i < len; i++)
	 {
	    int v = (int) sh.pixels [k * len + i];
	    acc += v * v;
	 }
	 sh.norm_i [nleaf - 1 + k] = acc;
      }
      __syncthreads ();
      for (int l = P.lmin + 1; l <= P.lc_max; l++)
      {
	 const unsigned nn    = 1u << (P.lc_max - l);
	 const unsigned node0 = nn - 1;
	 for (unsigned k = tid; k < nn; k += NT)
	    sh.norm_i [node0 + k] = sh.norm_i [2 * (node0 + k) + 1]
				    + sh.norm_i [2 * (node0 + k) + 2];
	 __syncthreads ();
      }
   }
   if (tid == 0 && g0 == 0)
   {
      unsigned ns = 0;
      /* need_image states: all states inside the image; count for the byte model */
      ns = band ? sh.h->y_states : sh.h->states;
      sh.h->blocks++;
      sh.h->ip_bytes += 4ull * size + 4ull * (63 + (unsigned) ((1 << (P.lc_max - P.il)) - 1)) * ns;
   }
   LAP_T (sh.h, LAP_PIX);
   cta_compute_T<NT, CLU> (P, W, sh, 0, 0, P.lc_max, P.lc_max, g0, gnt);
}

/* exact fp32 left-to-right sum of squares of a node (approx.c:388-389) */
__device__ float
t0_node_norm (const DevParams &P, const Sh &sh, unsigned image, unsigned address, int level)
{
   int s = sh.norm_i [image];

   if (s < 0)			/* the difference block of a nondeterministic prediction: cta_nd_range */
      return __uint_as_float ((unsigned) s & 0x7fffffffu);
   if (s <= (1 << 24))
      return (float) s;		/* every partial sum is an exactly representable integer */
   const float *px = sh.pixels + ((size_t) address << level);
   float	norm = 0;
   for (unsigned i = 0; i < (1u << level); i++)
      norm += px [i] * px [i];
   return norm;
}

/*****************************************************************************
	    new state: images and state x state rows  (codec/control.c, ip.c)
*****************************************************************************/

/* state x state product of states a, b at table level index li (symmetric storage) */
__device__ __forceinline__ float
ss_get (const DevParams &P, const TileWs &W, int li, unsigned a, unsigned b)
{
   return GP (W.SS) [((size_t) li * P.s_cap + a) * P.s_cap + b];
}

/* one entry <s, t> at level lmin + li for a state s whose images exist
   (ip.c:213-258 for levels > il, ip.c:297-323 below) */
__device__ float
dev_ss_entry (const DevParams &P, const TileWs &W, int li, unsigned s, unsigned t)
{
   const int level = P.lmin + li;

   if (level <= P.il)
   {
      const unsigned len = 1u << level;
      const float   *a	 = GP (W.img) + (size_t) s * FB_IMG_STRIDE + (len - 1);
      const float   *b	 = GP (W.img) + (size_t) t * FB_IMG_STRIDE + (len - 1);
      float	     ip	 = 0;
      for (unsigned i = 0; i < len; i++)
	 ip += a [i] * b [i];
      return ip;
   }
   float ip = 0;
   for (int label = 0; label < 2; label++)
   {
      const short *in1 = GP (W.into) + (size_t) (2 * s + label) * 6;
      const float *wt1 = GP (W.weight) + (size_t) (2 * s + label) * 6;
      const short *in2 = GP (W.into) + (size_t) (2 * t + label) * 6;
      const float *wt2 = GP (W.weight) + (size_t) (2 * t + label) * 6;
      const int	   c2  = GP (W.tree) [2 * t + label];
      int	   d1, d2;
      float	   sum;

      if ((d1 = GP (W.tree) [2 * s + label]) != FB_RANGE)
      {
	 sum = 0;
	 if (c2 != FB_RANGE)
	    sum = ss_get (P, W, li - 1, (unsigned) d1, (unsigned) c2);
	 for (int e2 = 0; (d2 = in2 [e2]) != FB_NO_EDGE; e2++)
	    sum += wt2 [e2] * ss_get (P, W, li - 1, (unsigned) d1, (unsigned) d2);
	 ip += sum;
      }
      for (int e1 = 0; (d1 = in1 [e1]) != FB_NO_EDGE; e1++)
      {
	 sum = 0;
	 if (c2 != FB_RANGE)
	    sum = ss_get (P, W, li - 1, (unsigned) d1, (unsigned) c2);
	 for (int e2 = 0; (d2 = in2 [e2]) != FB_NO_EDGE; e2++)
	    sum += wt2 [e2] * ss_get (P, W, li - 1, (unsigned) d1, (unsigned) d2);
	 ip += wt1 [e1] * sum;
      }
   }
   return ip;
}

/* codec/control.c:205-257 (compute_images) for one new state: every element of levels
   1..il is (child's element) + sum over edges of (domain element * weight) */
template <int NT>
__device__ void
cta_state_images (const DevParams &P, const TileWs &W, unsigned s)
{
   const int tid = threadIdx.x;
   const int tot = (1 << (P.il + 1)) - 1;

   for (int e = 1 + tid; e < tot; e += NT)
   {
      const int level = 31 - __clz (e + 1);	   /* e in [2^level - 1, 2^(level+1) - 1) */
      const int pos   = e - ((1 << level) - 1);
      const int half  = 1 << (level - 1);
      const int label = pos >= half;
      const int i     = pos - label * half;
      const int so    = (half - 1) + i;		   /* offset in the source image */
      int	dom   = GP (W.tree) [2 * s + label];
      float	v     = 0;

      if (dom != FB_RANGE)
	 v = GP (W.img) [(size_t) dom * FB_IMG_STRIDE + so];
      const short *in = GP (W.into) + (size_t) (2 * s + label) * 6;
      const float *wt = GP (W.weight) + (size_t) (2 * s + label) * 6;
      for (int k = 0; (dom = in [k]) != FB_NO_EDGE; k++)
	 v += GP (W.img) [(size_t) dom * FB_IMG_STRIDE + so] * wt [k];
      GP (W.img) [(size_t) s * FB_IMG_STRIDE + e] = v;
   }
   __syncthreads ();
}

/*
 *  codec/ip.c:184-260 for from == to == s.  Every term of the new row refers to states
 *  < s, so the levels do not depend on each other.  For a table level the needed rows
 *  <d1, .> of the level below (d1 = the new state's child / edge targets, a handful) are
 *  staged in shared memory with coalesced loads, then every partner state t gathers from
 *  them.  The summation order of ip.c:213-258 is kept: per label, per source of s, the
 *  inner sum over t's child and edges, scaled by the source's weight.
 */
template <int NT>
__device__ void
cta_state_products (const DevParams &P, const TileWs &W, const Sh &sh, unsigned s,
		    int first_li = 0, int step_li = 1)
{
   /* first_li, step_li: the table levels of this block when the blocks of a cluster share the
      new state's rows (the levels do not depend on each other) */
   const int tid  = threadIdx.x;
   ShHdr    *h	  = sh.h;
   /* scratch rows: the pursuit's work arrays are idle while a state is appended */
   /* scratch rows: the pursuit's work arrays (one contiguous area) are idle while a state is
      appended; rows of s + 1 floats are packed into it, so the shorter the rows the more
      source rows of the level below can be staged */
   const int stride = (int) ((s + 1 + 3) & ~3u);
   const int NR	    = min (2 * (FB_MAXEDGES + 1), sh.scratch_len / stride);
   unsigned  phase  = h->mbar_phase;	/* parity of the bulk copies' barrier (kept across calls) */

   if (tid == 0)
   {
      int nsrc = 0;

      for (int label = 0; label < 2; label++)
      {
	 int	      n	 = 0;
	 const int    c	 = GP (W.tree) [2 * s + label];
	 const short *in = GP (W.into) + (size_t) (2 * s + label) * 6;
	 const float *wp = GP (W.weight) + (size_t) (2 * s + label) * 6;

	 h->ap_child [label] = c != FB_RANGE;
	 if (c != FB_RANGE)
	 {
	    h->ap_dom [label][n] = (short) c;
	    h->ap_w [label][n]	 = 1.0f;
	    n++;
	 }
	 for (int e = 0; in [e] != FB_NO_EDGE && n < FB_MAXEDGES + 1; e++)
	 {
	    h->ap_dom [label][n] = in [e];
	    h->ap_w [label][n]	 = wp [e];
	    n++;
	 }
	 h->ap_cnt [label] = (unsigned char) n;
	 for (int k = 0; k < n; k++)
	 {
	    int j;
	    for (j = 0; j < nsrc; j++)
	       if (h->ap_src [j] == h->ap_dom [label][k])
		  break;
	    if (j == nsrc)
	       h->ap_src [nsrc++] = h->ap_dom [label][k];
	    h->ap_row [label][k] = (signed char) (j < NR ? j : -1);
	 }
      }
      h->ap_nsrc = nsrc;
   }
   __syncthreads ();

   for (int li = first_li; li < P.nlev; li += step_li)
   {
      const int level = P.lmin + li;

      if (level <= P.il)
      {
	 /* direct dot products of the state images (ip.c:297-323) */
	 const unsigned len = 1u << level;
	 const float   *a   = GP (W.img) + (size_t) s * FB_IMG_STRIDE + (len - 1);

	 for (unsigned t = tid; t <= s; t += NT)
	 {
	    if (!GP (W.domain_type) [t])
	       continue;
	    const float *b  = GP (W.img) + (size_t) t * FB_IMG_STRIDE + (len - 1);
	    float	 ip = 0;
	    if (level == 5)
	    {
	       /* the level-5 part is floats 31..62 of a 256-byte row: load the whole
		  upper half with aligned float4s (independent loads, one L2 round trip) */
	       const float4 *b4 = (const float4 *) (GP (W.img) + (size_t) t * FB_IMG_STRIDE + 32);
	       float4	     v [8];
	       float	     prev = GP (W.img) [(size_t) t * FB_IMG_STRIDE + 31];
#pragma unroll
	       for (int q = 0; q < 8; q++)
		  v [q] = b4 [q];
#pragma unroll
	       for (int q = 0; q < 8; q++)
	       {
		  ip += a [4 * q] * prev;
		  ip += a [4 * q + 1] * v [q].x;
		  ip += a [4 * q + 2] * v [q].y;
		  ip += a [4 * q + 3] * v [q].z;
		  prev = v [q].w;
	       }
	    }
	    else
	       for (unsigned i = 0; i < len; i++)
		  ip += a [i] * b [i];
	    GP (W.SS) [((size_t) li * P.s_cap + s) * P.s_cap + t] = ip;
	    GP (W.SS) [((size_t) li * P.s_cap + t) * P.s_cap + s] = ip;
	    if (t == s)
	       GP (W.diag) [(size_t) li * P.s_cap + s] = ip;
	 }
	 LAP (h, LAP_AP_DIRECT);
	 continue;
      }
      /* stage the source rows of the level below */
      const int nsrc = h->ap_nsrc < NR ? h->ap_nsrc : NR;
      /* rows are 16-byte aligned at both ends; the padded length stays inside the table row: one
	 bulk copy per row, issued by thread 0, all of them in flight together */
      if (tid == 0)
      {
#ifndef FB200_EMU	/* the scratch rows were last touched through the generic proxy (before a barrier) */
	 asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
#endif
	 mbar_expect (&h->mbar, (unsigned) (nsrc * stride * 4));
	 for (int j = 0; j < nsrc; j++)
	    bulk_copy_row (sh.num + (size_t) j * stride,
			   GP (W.SS) + ((size_t) (li - 1) * P.s_cap + h->ap_src [j]) * P.s_cap,
			   (unsigned) stride * 4, &h->mbar);
      }
      mbar_wait (&h->mbar, phase);
      phase ^= 1u;
#ifdef FB200_EMU
      __syncthreads ();		/* (the emulated copies are thread 0's own stores) */
#endif
      for (unsigned t = tid; t <= s; t += NT)
      {
	 if (!GP (W.domain_type) [t])
	    continue;
	 float	  ip = 0;
	 TransReg tr;

	 load_trans (GP (W.trans) + t, tr);
#pragma unroll
	 for (int label = 0; label < 2; label++)
	 {
	    const int c2 = tr.child [label];
	    const int n1 = h->ap_cnt [label];

	    for (int k = 0; k < n1; k++)
	    {
	       const int    j	= h->ap_row [label][k];
	       const float *row = j < 0 ? GP (W.SS) + ((size_t) (li - 1) * P.s_cap
						  + h->ap_dom [label][k]) * P.s_cap
					: sh.num + (size_t) j * stride;
	       float sum = 0;

	       if (c2 != FB_RANGE)
		  sum = row [c2];
#pragma unroll
	       for (int e2 = 0; e2 < FB_MAXEDGES; e2++)
		  if (tr.into [label][e2] != FB_NO_EDGE)
		     sum += tr.w [label][e2] * row [tr.into [label][e2]];
	       if (k == 0 && h->ap_child [label])
		  ip += sum;
	       else
		  ip += h->ap_w [label][k] * sum;
	    }
	 }
	 GP (W.SS) [((size_t) li * P.s_cap + s) * P.s_cap + t] = ip;
	 GP (W.SS) [((size_t) li * P.s_cap + t) * P.s_cap + s] = ip;
	 if (t == s)
	    GP (W.diag) [(size_t) li * P.s_cap + s] = ip;
      }
      __syncthreads ();
      LAP (h, LAP_AP_STAGED);
   }
   if (tid == 0)
      h->mbar_phase = phase;
}

/* codec/wfalib.c:154-180 */
__device__ float
t0_final_distribution (const TileWs &W, unsigned s)
{
   float final = 0;

   for (int label = 0; label < 2; label++)
   {
      int dom = GP (W.tree) [2 * s + label];

      if (dom != FB_RANGE)
	 final += GP (W.final_d) [dom];
      const short *in = GP (W.into) + (size_t) (2 * s + label) * 6;
      const float *wt = GP (W.weight) + (size_t) (2 * s + label) * 6;
      for (int e = 0; (dom = in [e]) != FB_NO_EDGE; e++)
	 final += wt [e] * GP (W.final_d) [dom];
   }
   return final / 2;
}

/* codec/wfalib.c:233-274 (edges kept sorted by target state) */
__device__ void
t0_append_edge (const TileWs &W, unsigned from, int into, float weight, int label)
{
   short *in = GP (W.into) + (size_t) (2 * from + label) * 6;
   float *wt = GP (W.weight) + (size_t) (2 * from + label) * 6;
   int	  pos, edge;

   for (pos = 0; in [pos] != FB_NO_EDGE && in [pos] < into; pos++)
      ;
   for (edge = pos; in [edge] != FB_NO_EDGE; edge++)
      ;
   for (edge++; edge != pos; edge--)
   {
      in [edge] = in [edge - 1];
      wt [edge] = wt [edge - 1];
   }
   in [edge] = (short) into;
   wt [edge] = weight;
}

template <int NT> __device__ void cta_cluster_post (const DevParams &P, const Sh &sh, bool with_models);

/*
 *  codec/control.c:48-131 (append_state).  The state's tree / edges are already in place.
 *  Must be called by all threads; 'auxiliary' and 'level' are uniform.
 */
template <int NT, bool CLU>
__device__ void
cta_append_state (const DevParams &P, const TileWs &W, const Sh &sh, int auxiliary,
		  int level_of_state)
{
   const unsigned s = sh.h->states;

   if (threadIdx.x == 0)
   {
      const float final = t0_final_distribution (W, s);

      GP (W.final_d) [s]	   = final;
      GP (W.level_of_state) [s] = (uint8_t) level_of_state;
      GP (W.domain_type) [s]	   = auxiliary ? 0 : 2;
      if (!auxiliary)
	 GP (W.img) [(size_t) s * FB_IMG_STRIDE] = final;
   }
   __syncthreads ();
   if (!auxiliary)
   {
      LAP (sh.h, LAP_DECIDE);
      cta_state_images<NT> (P, W, s);
      LAP (sh.h, LAP_AP_IMG);
      if (CLU && P.cluster > 1)
      {
	 /* the table levels are shared out over the blocks of the cluster */
	 if (threadIdx.x == 0)
	 {
	    sh.h->job.type   = CJ_APPEND;
	    sh.h->job.nested = 0;
	    sh.h->job.states = s;
	    sh.h->job.s	     = s;
	 }
	 cta_cluster_post<NT> (P, sh, false);
	 cta_state_products<NT> (P, W, sh, s, 0, P.cluster);
	 cl_sync ();
	 LAP (sh.h, LAP_CLUSTER);
      }
      else
	 cta_state_products<NT> (P, W, sh, s);
   }
   if (threadIdx.x == 0)
   {
      if (!auxiliary)
	 sh.h->ss_bytes += 4ull * (unsigned) P.nlev * (s + 1);
      sh.h->states = s + 1;
      if (s + 1 >= FB200_MAXSTATES)
	 sh.h->status = FB200_EMAXSTATES;
      else if (s + 1 >= (unsigned) P.s_cap)
	 sh.h->status = FB200_ECAPACITY;
   }
   __syncthreads ();
}

/* the built-in initial basis "small.fco" (input/basis.c:76-104,126-131) and
   append_basis_states (control.c:133-173) */
template <int NT>
__device__ void
cta_init_basis (const DevParams &P, const TileWs &W, const Sh &sh)
{
   const int tid = threadIdx.x;

   /* empty automaton (alloc_wfa, wfalib.c:98-113) for the states we may touch */
   for (unsigned i = tid; i < (unsigned) P.s_cap * 2; i += NT)
   {
      GP (W.tree) [i]     = FB_RANGE;
      GP (W.y_state) [i]  = FB_RANGE;
      GP (W.into) [(size_t) i * 6] = FB_NO_EDGE;
      GP (W.y_column) [i] = 0;
      GP (W.x) [i] = 0;
      GP (W.y) [i] = 0;
   }
   for (unsigned i = tid; i < (unsigned) P.s_cap; i += NT)
   {
      GP (W.domain_type) [i]	   = 0;
      GP (W.final_d) [i]	   = 0;
      GP (W.level_of_state) [i] = 0;
   }
   __syncthreads ();
   if (tid == 0)
   {
      GP (W.domain_type) [0] = 2;
      GP (W.final_d) [0]	= 128;
      t0_append_edge (W, 0, 0, 1.0f, 0);
      t0_append_edge (W, 0, 0, 1.0f, 1);
      GP (W.final_d) [1] = 64;
      GP (W.final_d) [2] = 64;
      GP (W.domain_type) [1] = 2;
      GP (W.domain_type) [2] = 2;
      t0_append_edge (W, 1, 2, 0.5f, 0);
      t0_append_edge (W, 1, 2, 0.5f, 1);
      t0_append_edge (W, 1, 0, 0.5f, 1);
      t0_append_edge (W, 2, 1, 1.0f, 0);
      t0_append_edge (W, 2, 1, 1.0f, 1);
      for (unsigned s = 0; s < 3; s++)
      {
	 t0_store_trans (W, s);
	 GP (W.img) [(size_t) s * FB_IMG_STRIDE] = GP (W.final_d) [s];
	 GP (W.level_of_state) [s]		    = (uint8_t) -1;
      }
      /* compute_images (0, 2): level outermost because the basis states refer to
	 each other */
      for (int level = 1; level <= P.il; level++)
	 for (unsigned s = 0; s < 3; s++)
	    for (int label = 0; label < 2; label++)
	    {
	       const int half = 1 << (level - 1);
	       float	*dst  = GP (W.img) + (size_t) s * FB_IMG_STRIDE + ((1 << level) - 1)
				+ label * half;
	       int	 dom  = GP (W.tree) [2 * s + label];

	       for (int i = 0; i < half; i++)
		  dst [i] = dom != FB_RANGE
			    ? GP (W.img) [(size_t) dom * FB_IMG_STRIDE + (half - 1) + i] : 0.0f;
	       const short *in = GP (W.into) + (size_t) (2 * s + label) * 6;
	       const float *wt = GP (W.weight) + (size_t) (2 * s + label) * 6;
	       for (int e = 0; (dom = in [e]) != FB_NO_EDGE; e++)
		  for (int i = 0; i < half; i++)
		     dst [i] += GP (W.img) [(size_t) dom * FB_IMG_STRIDE + (half - 1) + i] * wt [e];
	    }
      /* compute_ip_states_state (0, 2): level outermost */
      for (int li = 0; li < P.nlev; li++)
	 for (unsigned s1 = 0; s1 < 3; s1++)
	    for (unsigned s2 = 0; s2 <= s1; s2++)
	    {
	       const float ip = dev_ss_entry (P, W, li, s1, s2);

	       GP (W.SS) [((size_t) li * P.s_cap + s1) * P.s_cap + s2] = ip;
	       GP (W.SS) [((size_t) li * P.s_cap + s2) * P.s_cap + s1] = ip;
	       if (s1 == s2)
		  GP (W.diag) [(size_t) li * P.s_cap + s1] = ip;
	    }
      sh.h->states = 3;
   }
   __syncthreads ();
}

/*****************************************************************************
		     matching pursuit  (codec/approx.c:317-642)
*****************************************************************************/

__device__ __forceinline__ int
dom_state (const Sh &sh, const MpWork &w, int d)
{
   /* the y-state, when it is an extra domain, sits in slot pool_n (set per pursuit) */
   return (int) sh.pool [d];
}

/* per-step scalars every candidate needs (thread 0) */
__device__ void
t0_mp_prepare_step (const DevParams &P, const Sh &sh, MpRes &mp, int n)
{
   MpWork &w = sh.h->w;
   int	   nc = 0, ncs = 0;
   float   prefix = 0;

   w.n = n;
   for (int k = 0; k < n; k++)
   {
      w.fB [k] = w.B [k] / w.N [k];
      if (mp.weight [k] != 0)
      {
	 const int idx = mp.indices [k];
	 const int st  = dom_state (sh, w, idx);

	 w.cvec [nc++] = (short) idx;
	 if (st != w.y_state || w.y_state < 0)
	 {
	    /* insertion sort by domain index (rle_bits sorts with qsort) */
	    int p = ncs++;
	    while (p > 0 && w.csorted [p - 1] > idx)
	    {
	       w.csorted [p] = w.csorted [p - 1];
	       p--;
	    }
	    w.csorted [p] = (short) idx;
	 }
	 /* aac_bits prefix over the chosen weights (coeff.c:230-237) */
	 prefix = (float) ((double) prefix - (st ? w.l2_lv : w.l2_dc) [mp.code [k]]);
      }
   }
   w.nc	   = nc;
   w.ncs   = ncs;
   w.wb_nd = (float) ((double) prefix - w.l2_lv [w.half_lv]);
   w.wb_dc = (float) ((double) prefix - w.l2_dc [w.half_dc]);
}

/*
 *  Pass 1 for domain d (approx.c:433-462): rate of "chosen vectors + d with a dummy weight
 *  0.5", turned into the optimistic cost bound.  N = the largest step number of the
 *  configuration (max_elements - 1): ONE body serves every step of a pursuit, so that the
 *  second and third step find their code in the instruction cache.
 */
template <int N>
__device__ __forceinline__ float
mp_pass1 (const MpWork &w, int d, int st, float num, float den, float price, float err)
{
   short srt [N + 1];
   int	 ns = w.ncs;

#pragma unroll
   for (int e = 0; e < N + 1; e++)
      srt [e] = (e < N && e < ns) ? w.csorted [e] : (short) 0x7fff;
   if (d != w.ydom)
      sorted_insert<N + 1> (srt, ns, d);
   const float mb = dev_rle_bits<N + 1> (w.mbase, w.d0b, srt, ns, (unsigned) w.pool_n);
   const float wb = st ? w.wb_nd : w.wb_dc;

   return ((mb + wb + w.additional_bits) * price + err) - (num * num) / den;
}

/*
 *  Pass 2 for domain d at step n <= MAXN (approx.c:463-602): quantised weights by back
 *  substitution, true rate, true error.  Returns the costs; weights / bits / err in out[].
 *  out: [0..4] weights, [5] matrix bits, [6] weights bits, [7] err.  One body for all steps
 *  (see mp_pass1); the loops are unrolled to MAXN and guarded by the (uniform) step number.
 */
template <int MAXN>
__device__ __forceinline__ float
mp_pass2 (const DevParams &P, const Sh &sh, const MpWork &w, const MpRes &mp, int n, int d,
	  float num, float den, float price, float *out, int *cod)
{
   const int dcap = sh.dcap;
   float     f [MAXN + 1], r [MAXN + 1];
   int	     v [MAXN + 1], c [MAXN + 1];
   bool	     nd [MAXN + 1];	/* not the DC domain: ordinary quantiser */

#pragma unroll
   for (int k = 0; k <= MAXN; k++)
   {
      f [k]  = k < n ? w.fB [k] : num / den;
      v [k]  = k < n ? (int) mp.indices [k] : d;
      c [k]  = -1;
      r [k]  = 0;
      nd [k] = false;
   }
#pragma unroll
   for (int l = MAXN; l >= 0; l--)
      if (l <= n)
      {
	 /* btor (rtob (x)): the code, then its value from the table (rtob o btor is the
	    identity on codes, tests/test_oracle_golden.py) */
	 nd [l] = dom_state (sh, w, v [l]) != 0;
	 c [l]	= dev_rtob (f [l], nd [l] ? P.rpf_m : P.dc_m, nd [l] ? P.rpf_range : P.dc_range);
	 const float q = c [l] < 0 ? 0.0f : (nd [l] ? sh.qt_lv : sh.qt_dc) [c [l]];
	 f [l] = q;
	 r [l] = q;
#pragma unroll
	 for (int k = 0; k < l; k++)
	    f [k] -= q * sh.G [k * dcap + v [l]] / w.N [k];
      }
   /* rate of the quantised combination */
   float w_bits = 0, m_bits;
   {
      short srt [MAXN + 1];
      int   cnt = 0;

#pragma unroll
      for (int e = 0; e < MAXN + 1; e++)
	 srt [e] = (short) 0x7fff;
#pragma unroll
      for (int k = 0; k <= MAXN; k++)
	 if (k <= n && c [k] >= 0)
	 {
	    w_bits = (float) ((double) w_bits - (nd [k] ? w.l2_lv : w.l2_dc) [c [k]]);
	    if (v [k] != w.ydom)
	       sorted_insert<MAXN + 1> (srt, cnt, v [k]);
	 }
      m_bits = dev_rle_bits<MAXN + 1> (w.mbase, w.d0b, srt, cnt, (unsigned) w.pool_n);
   }
   /* back to the orthogonal basis, error (approx.c:571-586) */
#pragma unroll
   for (int k = 0; k < MAXN; k++)
#pragma unroll
      for (int l = k + 1; l <= MAXN; l++)
	 if (l <= n)
	    r [k] += sh.G [k * dcap + v [l]] * r [l] / w.N [k];
   float m_err = w.norm;
#pragma unroll
   for (int k = 0; k <= MAXN; k++)
      if (k <= n)
      {
	 const float Nk = k == n ? den : w.N [k];
	 const float Bk = k == n ? num : w.B [k];
	 m_err += (r [k] * r [k]) * Nk - 2 * r [k] * Bk;
      }
#pragma unroll
   for (int k = 0; k < FB_MAXEDGES; k++)
   {
      out [k] = (k <= MAXN && k <= n) ? f [k < MAXN ? k : MAXN] : 0.0f;
      cod [k] = (k <= MAXN && k <= n) ? c [k < MAXN ? k : MAXN] : -1;
   }
   out [5] = m_bits;
   out [6] = w_bits;
   out [7] = m_err;
   return (m_bits + w_bits + w.additional_bits) * price + m_err;
}

/*
 *  Step n (<= N) of the pursuit: find the domain the reference's index-ordered scan with its
 *  running minimum (approx.c:420-603) would select.
 *
 *  Phase 1: every thread computes the pass-1 bound of its domains.
 *  Waves:   the (few) domains whose bound beats the running minimum are evaluated lazily
 *	     in index order, 32 at a time, by warp 0; after each wave the minimum has
 *	     dropped and most remaining domains no longer qualify -- exactly the set the
 *	     sequential scan would have evaluated, plus at most 31 speculative ones per wave.
 *  Result in w.index / w.best_* / w.min_costs.  Ends with a barrier.
 */
template <int NT, int N>
__device__ void
cta_mp_find (const DevParams &P, const Sh &sh, MpRes &mp, float price, int n)
{
   const int tid  = threadIdx.x;
   const int lane = tid & 31;
   const int warp = tid >> 5;
   MpWork   &w	  = sh.h->w;
   const int D	  = w.D;
   const int D32  = (D + 31) & ~31;
   const float err = mp.err;
   float       m   = w.min_costs;

   for (int d = tid; d < D32; d += NT)
   {
      float b = INFINITY;

      if (d < D && !sh.used [d])
      {
	 b = mp_pass1<N> (w, d, dom_state (sh, w, d), sh.num [d], sh.den [d], price, err);
      }
      sh.bnd [d] = b;
      /* candidates of the first wave (a warp covers one 32-domain word) */
      const unsigned mask = __ballot_sync (0xffffffffu, b < m);
      if (lane == 0)
	 sh.cmask [d >> 5] = mask;
   }
   LAP (sh.h, LAP_MP_P1);
   int	pos   = 0;
   bool first = true;

   for (;;)
   {
      if (!first)
	 for (int d = tid; d < D32; d += NT)
	 {
	    const unsigned mask = __ballot_sync (0xffffffffu, d >= pos && sh.bnd [d] < m);
	    if (lane == 0)
	       sh.cmask [d >> 5] = mask;
	 }
      __syncthreads ();
      if (warp == 0)
      {
	 const int nwords = D32 >> 5;
	 int	   taken  = 0, cand = -1;
	 bool	   more	  = false;

	 for (int g = 0; g < nwords && !more; g += 32)
	 {
	    const unsigned word = (g + lane < nwords) ? sh.cmask [g + lane] : 0u;
	    const int	   cnt	= __popc (word);
	    int		   incl = cnt;
#pragma unroll
	    for (int o = 1; o < 32; o <<= 1)
	    {
	       const int t = __shfl_up_sync (0xffffffffu, incl, o);
	       if (lane >= o)
		  incl += t;
	    }
	    const int total = __shfl_sync (0xffffffffu, incl, 31);
	    /* lane j takes candidate number j - taken of this group of words: the word is
	       the first whose inclusive prefix count exceeds that rank */
	    const int r	 = lane - taken;
	    int	      lo = 0, hi = 31;
#pragma unroll
	    for (int it = 0; it < 5; it++)
	    {
	       const int mid = (lo + hi) >> 1;
	       const int v   = __shfl_sync (0xffffffffu, incl, mid);
	       if (v > r)
		  hi = mid;
	       else
		  lo = mid + 1;
	    }
	    const unsigned wsel = __shfl_sync (0xffffffffu, word, lo);
	    const int	   esel = __shfl_sync (0xffffffffu, incl - cnt, lo);
	    if (r >= 0 && r < total)
	       cand = ((g + lo) << 5) + (int) __fns (wsel, 0, r - esel + 1);
	    if (taken + total > 32)
	       more = true;
	    taken = taken + total > 32 ? 32 : taken + total;
	    if (taken == 32 && g + 32 < nwords)
	       more = true;	/* unscanned words may hold further candidates */
	 }
	 float	   key = INFINITY, costs = 0;
	 float	   res [8];
	 int	   cod [FB_MAXEDGES];
	 const int d = cand;

	 if (lane < taken)
	 {
	    const float b = sh.bnd [d];
	    costs = mp_pass2<N> (P, sh, w, mp, n, d, sh.num [d], sh.den [d], price, res, cod);
	    key	  = b > costs ? b : costs;	/* both must beat the running minimum */
	 }
	 /* ordered resolution (lanes are in index order) */
	 int winlane = -1, last = -1;
	 for (;;)
	 {
	    const unsigned m2 = __ballot_sync (0xffffffffu, key < m && lane > last);
	    if (!m2)
	       break;
	    last    = __ffs ((int) m2) - 1;
	    m	    = __shfl_sync (0xffffffffu, costs, last);
	    winlane = last;
	 }
	 if (lane == winlane)
	 {
#pragma unroll
	    for (int k = 0; k < FB_MAXEDGES; k++)
	    {
	       w.best_f [k] = res [k];
	       w.best_c [k] = (short) cod [k];
	    }
	    w.best_mbits = res [5];
	    w.best_wbits = res [6];
	    w.best_err	 = res [7];
	    w.best_costs = m;
	    w.min_costs	 = m;
	    w.index	 = d;
	 }
	 const int last_cand = __shfl_sync (0xffffffffu, cand, 31);
	 if (lane == 0)
	 {
	    if (first && winlane < 0)
	       w.index = -1;
	    w.wave_done = !more;
	    w.wave_pos	= more ? last_cand + 1 : D;
	    if (taken)
	       w.st_pass2 += (unsigned) taken;
	 }
      }
      __syncthreads ();
      if (w.wave_done)
	 break;
      m	    = w.min_costs;
      pos   = w.wave_pos;
      first = false;
   }
}

/*
 *  The same step for the shape that has an SM to itself: ALL candidates whose bound beats the
 *  minimum the step starts with (up to one per thread) are evaluated at once, every warp its 32;
 *  warp 0 then resolves them in index order against the running minimum exactly as above.  More
 *  speculative evaluations, but (almost always) a single wave: the narrow version needs 1.5 waves
 *  per step on a 1024^2 frame, each a dependent chain of ~2500 cycles on the path of the stream.
 *  A candidate's result travels through shared memory (sh.wide, one column per thread).
 */
template <int NT, int N>
__device__ void
cta_mp_find_wide (const DevParams &P, const Sh &sh, MpRes &mp, float price, int n)
{
   const int tid  = threadIdx.x;
   const int lane = tid & 31;
   const int warp = tid >> 5;
   MpWork   &w	  = sh.h->w;
   const int D	  = w.D;
   const int D32  = (D + 31) & ~31;
   const float err = mp.err;
   float       m   = w.min_costs;
   float      *ck  = sh.wide, *cc = sh.wide + NT, *pay = sh.wide + 2 * NT;

   for (int d = tid; d < D32; d += NT)
   {
      float b = INFINITY;

      if (d < D && !sh.used [d])
	 b = mp_pass1<N> (w, d, dom_state (sh, w, d), sh.num [d], sh.den [d], price, err);
      sh.bnd [d] = b;
      const unsigned mask = __ballot_sync (0xffffffffu, b < m);
      if (lane == 0)
	 sh.cmask [d >> 5] = mask;
   }
   LAP (sh.h, LAP_MP_P1);
   int	pos   = 0;
   bool first = true;

   for (;;)
   {
      if (!first)
	 for (int d = tid; d < D32; d += NT)
	 {
	    const unsigned mask = __ballot_sync (0xffffffffu, d >= pos && sh.bnd [d] < m);
	    if (lane == 0)
	       sh.cmask [d >> 5] = mask;
	 }
      __syncthreads ();
      /* every warp scans the candidate words (the same counts in every warp) and picks the
	 candidates number 32 * warp ... 32 * warp + 31 */
      const int nwords = D32 >> 5;
      int	base = 0, cand = -1;
      bool	more = false;

      for (int g = 0; g < nwords; g += 32)
      {
	 const unsigned word = (g + lane < nwords) ? sh.cmask [g + lane] : 0u;
	 const int	cnt  = __popc (word);
	 int		incl = cnt;
#pragma unroll
	 for (int o = 1; o < 32; o <<= 1)
	 {
	    const int t = __shfl_up_sync (0xffffffffu, incl, o);
	    if (lane >= o)
	       incl += t;
	 }
	 const int total = __shfl_sync (0xffffffffu, incl, 31);
	 const int r	 = tid - base;
	 int	   lo = 0, hi = 31;
#pragma unroll
	 for (int it = 0; it < 5; it++)
	 {
	    const int mid = (lo + hi) >> 1;
	    const int v	  = __shfl_sync (0xffffffffu, incl, mid);
	    if (v > r)
	       hi = mid;
	    else
	       lo = mid + 1;
	 }
	 const unsigned wsel = __shfl_sync (0xffffffffu, word, lo);
	 const int	esel = __shfl_sync (0xffffffffu, incl - cnt, lo);
	 if (r >= 0 && r < total)
	    cand = ((g + lo) << 5) + (int) __fns (wsel, 0, r - esel + 1);
	 base += total;
	 if (base >= NT && g + 32 < nwords)
	 {
	    more = true;		/* unscanned words may hold further candidates */
	    break;
	 }
      }
      if (base > NT)
	 more = true;
      const int ntake = base < NT ? base : NT;
      LAP_W (sh.h, LAP_PIX);		/* barrier + scan */

      if (tid < ntake)
      {
	 float	     res [8];
	 int	     cod [FB_MAXEDGES];
	 const float b	   = sh.bnd [cand];
	 const float costs = mp_pass2<N> (P, sh, w, mp, n, cand, sh.num [cand], sh.den [cand], price, res, cod);

	 ck [tid] = b > costs ? b : costs;	/* both must beat the running minimum */
	 cc [tid] = costs;
#pragma unroll
	 for (int k = 0; k < FB_MAXEDGES; k++)
	 {
	    pay [k * NT + tid]			= res [k];
	    pay [(FB_MAXEDGES + k) * NT + tid] = __int_as_float (cod [k]);
	 }
	 pay [10 * NT + tid] = res [5];
	 pay [11 * NT + tid] = res [6];
	 pay [12 * NT + tid] = res [7];
	 if (more && tid == ntake - 1)
	    w.wave_pos = cand + 1;
      }
      LAP_W (sh.h, LAP_DOTS);		/* thread 0's own candidate */
      __syncthreads ();
      LAP_W (sh.h, LAP_UPSWEEP);	/* waiting for the other candidates */
      if (warp == 0)
      {
	 int win = -1;

	 for (int c = 0; c < ntake; c += 32)
	 {
	    const float key   = c + lane < ntake ? ck [c + lane] : INFINITY;
	    const float costs = c + lane < ntake ? cc [c + lane] : 0.0f;
	    int		last  = -1;

	    for (;;)
	    {
	       const unsigned m2 = __ballot_sync (0xffffffffu, key < m && lane > last);
	       if (!m2)
		  break;
	       last = __ffs ((int) m2) - 1;
	       m    = __shfl_sync (0xffffffffu, costs, last);
	       win  = c + last;
	    }
	 }
	 if (lane == 0)
	 {
	    w.win = win;
	    if (win >= 0)
	    {
#pragma unroll
	       for (int k = 0; k < FB_MAXEDGES; k++)
	       {
		  w.best_f [k] = pay [k * NT + win];
		  w.best_c [k] = (short) __float_as_int (pay [(FB_MAXEDGES + k) * NT + win]);
	       }
	       w.best_mbits = pay [10 * NT + win];
	       w.best_wbits = pay [11 * NT + win];
	       w.best_err   = pay [12 * NT + win];
	       w.best_costs = m;
	       w.min_costs  = m;
	    }
	    else if (first)
	       w.index = -1;
	    w.wave_done = !more;
	    if (!more)
	       w.wave_pos = D;
	    w.st_pass2 += (unsigned) ntake;
	 }
      }
      __syncthreads ();
      if (tid == w.win)
	 w.index = cand;
      __syncthreads ();
      if (w.wave_done)
	 break;
      m	    = w.min_costs;
      pos   = w.wave_pos;
      first = false;
   }
}

/*
 *  One matching pursuit over the current pool for the range (level, image) of the block's
 *  product tree.  'mp' lives in shared memory; the work counters of the call are left in
 *  w.st_* / w.D for whoever uses the result; 'excluded' is the domain index the second_domain_block retry
 *  must not use (approx.c:112-116), or -1.
 */
template <int NT>
__device__ void
cta_matching_pursuit (const DevParams &P, const TileWs &W, const Sh &sh, MpRes &mp,
		      int level, unsigned image, float norm, float tree_bits,
		      float mv_tree_bits, float price, int y_state_in, int excluded)
{
   /* norm: the squared norm of the range (t0_node_norm), needed by thread 0 only */
   const int	tid	 = threadIdx.x;
   MpWork      &w	 = sh.h->w;
   const float	min_norm = 2e-3f;
   const int	dcap	 = sh.dcap;

   LAP (sh.h, LAP_ENTER);
   /* ---- prologue: per-call tables ---- */
   if (tid == 0)
   {
      int y_state = y_state_in;

      if (y_state >= 0 && !(GP (W.domain_type) [y_state] & 2))
	 y_state = -1;
      w.y_state = y_state;
      w.pool_n	= BLOB_U16 (sh, MB_N);
      w.ydom	= -1;
      w.D	= w.pool_n;
      if (y_state >= 0)
      {
	 /* rle_generate (domain-pool.c:707-735): the y-state is an extra domain unless
	    it already is a pool member */
	 int member = -1;
	 for (int d = 0; d < w.pool_n; d++)
	    if (sh.pool [d] == y_state)
	       member = d;
	 if (member >= 0)
	    w.ydom = member;
	 else
	 {
	    w.ydom = w.pool_n;
	    w.D	   = w.pool_n + 1;
	    sh.pool [w.pool_n] = (short) y_state;
	    if (w.pool_n < sh.h->pool_lo)
	       sh.h->pool_lo = w.pool_n;
	 }
      }
      w.level = level;
      w.li    = level - P.lmin;
      w.size  = 1 << level;
      w.price = price;
      w.norm  = norm;
      w.additional_bits = tree_bits + mv_tree_bits + 0.0f + 0.0f + 0.0f;
      w.d0b [0] = t0_d0_bits (sh, 0, y_state, c_matrix_0, c_matrix_1);
      w.d0b [1] = t0_d0_bits (sh, 1, y_state, c_matrix_0, c_matrix_1);
      w.st_steps = w.st_pass2 = 0;
   }
   /*
    *  Grey bands (no y-state): nothing the other threads need depends on thread 0's scalars,
    *  so the tables and the gathers below proceed next to them; with a y-state the pool
    *  size and the y-slot come from thread 0 first.
    */
   const bool grey = y_state_in < 0;	/* uniform */

   if (!grey)
      __syncthreads ();
   /* log2 tables of the models (the models do not change during one pursuit); entries are
      handed out from the last thread down: warp 0 is busy with the scalars */
   {
      const int	   ctx	  = level - P.coeff_min_level;
      const short *counts = sh.blob + MB_COUNTS;
      const short *lv	  = counts + P.aac_dc_size + ctx * P.aac_lvl_size;

      for (int i = NT - 1 - tid; i < P.aac_dc_size + P.aac_lvl_size + FB_MAXEDGES + 2; i += NT)
      {
	 if (i < P.aac_dc_size)
	    w.l2_dc [i] = dev_log2d (counts [i] / (float) sh.blob [MB_TOTALS]);
	 else if (i < P.aac_dc_size + P.aac_lvl_size)
	 {
	    const int c = i - P.aac_dc_size;
	    w.l2_lv [c] = dev_log2d (lv [c] / (float) sh.blob [MB_TOTALS + ctx + 1]);
	 }
	 else
	 {
	    const int k = i - P.aac_dc_size - P.aac_lvl_size;
	    w.mbase [k] = neg_log2f_via_double (BLOB_S16 (sh, MB_COUNT + (k <= FB_MAXEDGES ? k : FB_MAXEDGES))
						/ (float) BLOB_U16 (sh, MB_TOTAL));
	 }
      }
   }
   const int   D     = grey ? (int) BLOB_U16 (sh, MB_N) : w.D;
   const int   li    = level - P.lmin;
   const float fsize = (float) (1 << level);

   /* ---- numerators / denominators (approx.c:358-374) ---- */
   /* every thread's gathers are in flight together (cp.async, 4 bytes each), then the tests
      run on the thread's own entries */
   for (int d = tid; d < D; d += NT)
   {
      const int st = dom_state (sh, w, d);

      gather_f32_async (sh.den + d, GP (W.diag) + (size_t) li * P.s_cap + st);
      gather_f32_async (sh.num + d, GP (W.T) + (size_t) image * P.s_cap + st);
   }
   async_wait_all ();
   for (int d = tid; d < D; d += NT)
   {
      unsigned char used = 0;

      if (sh.den [d] / fsize < min_norm)
	 used = 1;
      else if (fabsf (sh.num [d]) < min_norm)
	 used = 1;
      if (d == excluded)
	 used = 1;
      sh.used [d] = used;
   }
   __syncthreads ();
   if (tid == 0)
   {
      /* costs of the empty linear combination (approx.c:391-400) */
      mp.err	      = w.norm;
      mp.weights_bits = 0;
      mp.matrix_bits  = dev_rle_bits<1> (w.mbase, w.d0b, w.csorted, 0, (unsigned) w.pool_n);
      mp.costs	      = (mp.matrix_bits + mp.weights_bits + w.additional_bits) * price
			+ mp.err;
      t0_mp_prepare_step (P, sh, mp, 0);
      w.min_costs = mp.costs;
   }
   __syncthreads ();
   LAP (sh.h, LAP_MP_PRO);

   int n = 0;
   for (;;)
   {
      if (NT >= 512)		/* (the shape with an SM to itself) */
      {
	 if (P.max_elements <= 3)
	    cta_mp_find_wide<NT, 2> (P, sh, mp, price, n);
	 else
	    cta_mp_find_wide<NT, 4> (P, sh, mp, price, n);
      }
      else if (P.max_elements <= 3)
	 cta_mp_find<NT, 2> (P, sh, mp, price, n);
      else
	 cta_mp_find<NT, 4> (P, sh, mp, price, n);

      /* ---- commit the step (approx.c:605-632) ---- */
      const int index = w.index;
      LAP (sh.h, LAP_MP_WAVES);
      if (index < 0)
	 break;
      if (tid == 0)
      {
	 /* not full_search: a found vector always improves the total */
	 mp.costs	 = w.best_costs;
	 mp.err		 = w.best_err;
	 mp.matrix_bits	 = w.best_mbits;
	 mp.weights_bits = w.best_wbits;
#pragma unroll 1
	 for (int k = 0; k <= n; k++)
	 {
	    mp.weight [k] = w.best_f [k];
	    mp.code [k]	  = w.best_c [k];
	 }
	 mp.indices [n] = (short) index;
	 mp.into [n]	= (short) dom_state (sh, w, index);
	 sh.used [index] = 1;
	 w.B [n] = sh.num [index];
	 w.N [n] = sh.den [index];
	 w.st_steps++;
	 if (n + 1 < P.max_elements)
	    t0_mp_prepare_step (P, sh, mp, n + 1);
      }
      __syncthreads ();
      LAP (sh.h, LAP_MP_COMMIT);
      if (n + 1 >= P.max_elements)
      {
	 n++;
	 break;
      }
      /* ---- Gram-Schmidt step n over the pool (approx.c:644-699) ---- */
      {
	 const int   sidx = dom_state (sh, w, index);
	 const float Nn	  = w.N [n], Bn = w.B [n];
	 const float *row = GP (W.SS) + ((size_t) li * P.s_cap + sidx) * P.s_cap;

	 /* the row entries of the thread's domains travel together into the (idle) bound array */
	 for (int d = tid; d < D; d += NT)
	    if (!sh.used [d])
	       gather_f32_async (sh.bnd + d, row + dom_state (sh, w, d));
	 async_wait_all ();
	 for (int d = tid; d < D; d += NT)
	    if (!sh.used [d])
	    {
	       float tmp = sh.bnd [d];

#pragma unroll 1
	       for (int k = 0; k < n; k++)
		  tmp -= sh.G [k * dcap + d] / w.N [k] * sh.G [k * dcap + index];
	       sh.G [n * dcap + d] = tmp;
	       const float den = sh.den [d] - (tmp * tmp) / Nn;
	       sh.den [d] = den;
	       sh.num [d] = sh.num [d] - Bn / Nn * tmp;
	       if (den / fsize < min_norm)
		  sh.used [d] = 1;
	    }
      }
      LAP (sh.h, LAP_MP_ORTHO);
      n++;
      /* no barrier needed here: the next pass touches only the thread's own domains
	 (same tid -> same d) plus data published before the last barrier */
   }

   if (tid == 0)
   {
      mp.indices [n] = FB_NO_EDGE;	/* best_n == n without full_search */
      mp.costs = (mp.matrix_bits + mp.weights_bits + w.additional_bits) * price + mp.err;
   }
   __syncthreads ();
}

/* rle_update (domain-pool.c:795-830) + inlined qac_update of the DC model (:404-446) */
__device__ void
t0_rle_update (const Sh &sh, const short *into, int y_state)
{
   /* into: the states of the used domains (what the reference looks up through the pool list) */
   int	    state_0 = 0, state_y = 0, edge = 0;

   for (edge = 0; into [edge] != FB_NO_EDGE; edge++)
   {
      const int st = into [edge];

      if (st == 0)
	 state_0 = 1;
      else if (st == y_state)
	 state_y = 1;
   }
   BLOB_S16 (sh, MB_COUNT + edge)++;
   BLOB_U16 (sh, MB_TOTAL)++;
   {
      int y_is_domain = 0, used_y = 0;

      if (BLOB_U16 (sh, MB_D0N) > 0)
      {
	 BLOB_S16 (sh, MB_D0INDEX)++;
	 if (0 == y_state)
	    y_is_domain = 1;
      }
      if (state_0)
      {
	 if (0 == y_state)
	 {
	    if (y_is_domain)
	       BLOB_S16 (sh, MB_D0INDEX)--;
	    BLOB_U16 (sh, MB_D0YINDEX) >>= 1;
	    used_y = 1;
	 }
	 else
	 {
	    BLOB_S16 (sh, MB_D0INDEX)--;
	    BLOB_S16 (sh, MB_D0INDEX) >>= 1;
	 }
      }
      if (y_state >= 0 && !used_y)
	 BLOB_U16 (sh, MB_D0YINDEX)++;
      if (BLOB_U16 (sh, MB_D0N) > 0 && BLOB_S16 (sh, MB_D0INDEX) > 1020)
	 BLOB_S16 (sh, MB_D0INDEX) = 1020;
      if (BLOB_U16 (sh, MB_D0YINDEX) > 1020)
	 BLOB_U16 (sh, MB_D0YINDEX) = 1020;
   }
   if (state_y)
      BLOB_U16 (sh, MB_YINDEX) >>= 1;
   else
      BLOB_U16 (sh, MB_YINDEX)++;
   if (BLOB_U16 (sh, MB_YINDEX) > 1020)
      BLOB_U16 (sh, MB_YINDEX) = 1020;
}

/* coeff.c:242-267 */
__device__ void
t0_aac_update (const DevParams &P, const Sh &sh, const short *code, const short *into,
	       int level)
{
   const int ctx    = level - P.coeff_min_level;
   short    *counts = sh.blob + MB_COUNTS;
   short    *lv	    = counts + P.aac_dc_size + ctx * P.aac_lvl_size;

   for (int e = 0; into [e] != FB_NO_EDGE; e++)
      if (into [e])
      {
	 lv [code [e]]++;
	 sh.blob [MB_TOTALS + ctx + 1]++;
      }
      else
      {
	 counts [code [e]]++;
	 sh.blob [MB_TOTALS]++;
      }
}

/*
 *  Thread 0: approximate_range after its pursuit(s) (approx.c:208-271) for the result 'mp': accept it
 *  if it is cheaper than max_costs, drop zero weights, fill the range.  lazy == NULL: the models take the
 *  range at once (rle_update, aac_update) -- a range that cannot be subdivided.  Otherwise the codes and
 *  the y-state go into the activation record 'lazy' and the models are updated if and when the linear
 *  combination wins (ST_DECIDE): the reference updates at once, keeps the result aside as "lc models" and
 *  restores the models it had (subdivide.c:226-237) -- the same thing, one copy of the models later.
 *  Returns the costs (MAXCOSTS: rejected), also left in h->ret_costs.
 */
__device__ float
t0_range_epilogue (const DevParams &P, const TileWs &W, const Sh &sh, MpRes &mp, int eff_ystate,
		   float max_costs, float price, int y_state, RangeRes *out, int level, unsigned image,
		   unsigned address, unsigned x, unsigned y, Frame *lazy)
{
   ShHdr *h = sh.h;
   float  costs;
   int	  n_edges = -1;

   if (mp.costs < max_costs)
   {
      int new_index = 0, edge;

      for (int old = 0; mp.indices [old] != FB_NO_EDGE; old++)
	 if (mp.weight [old] != 0)
	 {
	    mp.indices [new_index] = mp.indices [old];
	    mp.into [new_index]	   = mp.into [old];
	    mp.weight [new_index]  = mp.weight [old];
	    mp.code [new_index]	   = mp.code [old];
	    new_index++;
	 }
      mp.indices [new_index] = FB_NO_EDGE;
      mp.into [new_index]    = FB_NO_EDGE;
      if (!lazy)
      {
	 t0_rle_update (sh, mp.into, eff_ystate);
	 t0_aac_update (P, sh, mp.code, mp.into, level);
      }
      else
      {
	 for (edge = 0; edge < FB_MAXEDGES; edge++)
	    lazy->lc_code [edge] = mp.code [edge];
	 lazy->lc_ystate = (short) eff_ystate;
      }
      for (edge = 0; mp.indices [edge] != FB_NO_EDGE; edge++)
      {
	 out->into [edge]   = mp.into [edge];
	 out->weight [edge] = mp.weight [edge];
      }
      out->into [edge]	= FB_NO_EDGE;
      out->matrix_bits	= mp.matrix_bits;
      out->weights_bits = mp.weights_bits;
      out->err		= mp.err;
      costs		= mp.costs;
      n_edges		= edge;
   }
   else
   {
      out->into [0] = FB_NO_EDGE;
      costs	    = FB_MAXCOSTS;
   }
   h->ret_costs = costs;
   if (GP (W.trace) && h->trace_len < P.trace_cap)
   {
      fb200_trace_rec_t *t = GP (W.trace) + h->trace_len;

      t->level	 = (uint16_t) level;
      t->image	 = (uint16_t) image;
      t->address = (uint16_t) address;
      t->x	 = (uint16_t) x;
      t->y	 = (uint16_t) y;
      t->y_state = (int16_t) y_state;
      t->states	 = (uint16_t) h->states;
      t->n_edges = (int16_t) n_edges;
      t->max_costs    = max_costs;
      t->price	      = price;
      t->costs	      = costs;
      t->err	      = n_edges >= 0 ? out->err : 0;
      t->matrix_bits  = n_edges >= 0 ? out->matrix_bits : 0;
      t->weights_bits = n_edges >= 0 ? out->weights_bits : 0;
      for (int e = 0; e < 6; e++)
      {
	 t->into [e]   = e < n_edges ? out->into [e] : (int16_t) -1;
	 t->weight [e] = e < n_edges ? out->weight [e] : 0;
      }
   }
   if (GP (W.trace))
      h->trace_len++;
   return costs;
}

/*
 *  codec/approx.c:74-271 (approximate_range) for the still-image option set
 *  (second_domain_block optional).  Result in 'out' / return value in sh.h->ret_costs.
 */
template <int NT>
__device__ void
cta_approximate_range (const DevParams &P, const TileWs &W, const Sh &sh, float max_costs,
		       float price, int y_state, RangeRes *out, int level, unsigned image,
			       unsigned address, unsigned x, unsigned y, float mv_tree_bits,
		       Frame *lazy, int spec_k = 0)
{
   ShHdr *h = sh.h;

   /* one call site (one copy of the pursuit's code); the second round is the
      second_domain_block retry without the first domain (approx.c:103-127) */
   if (spec_k > 0)
   {
      /* the pursuit of this range ran ahead, next to the one of the range at the head of the
	 spine: same models, same states (cluster per stream) */
      if (threadIdx.x == 0)
      {
	 const SpecRes &r = h->spec [spec_k];

	 h->mp	      = r.mp;
	 h->w.y_state = -1;
	 h->mp_calls++;
	 h->mp_steps += r.steps;
	 h->pass2    += r.pass2;
	 h->mp_bytes += 8ull * r.D + 4ull * r.D * r.steps;
      }
      __syncthreads ();
   }
   else
   for (int round = 0; round <= (P.second_domain_block ? 1 : 0); round++)
   {
      MpRes &m = round ? h->tmp : h->mp;

      if (round)
      {
	 if (threadIdx.x == 0)
	    h->tmp = h->mp;
	 __syncthreads ();
      }
      /* (the prologue of the pursuit starts with thread-0 work followed by a barrier) */
      cta_matching_pursuit<NT> (P, W, sh, m, level, image,
				threadIdx.x == 0 ? t0_node_norm (P, sh, image, address, level) : 0.0f,
				out->tree_bits, mv_tree_bits, price, y_state,
				round ? (int) h->mp.indices [0] : -1);
      if (threadIdx.x == 0)
      {
	 const MpWork &w = h->w;

	 h->mp_calls++;
	 h->mp_steps += w.st_steps;
	 h->pass2    += w.st_pass2;
	 h->mp_bytes += 8ull * (unsigned) w.D + 4ull * (unsigned) w.D * w.st_steps;
      }
   }
   if (P.second_domain_block)
   {
      if (threadIdx.x == 0 && h->tmp.costs < h->mp.costs)
	 h->mp = h->tmp;
      __syncthreads ();
   }
   if (threadIdx.x == 0)
      t0_range_epilogue (P, W, sh, h->mp, h->w.y_state, max_costs, price, y_state, out, level, image, address,
			 x, y, lazy);
   __syncthreads ();
   LAP (h, LAP_AR_EPI);
}

/*****************************************************************************
		     model snapshots (subdivide.c:188-237)
*****************************************************************************/

template <int NT>
__device__ void
cta_copy_s16 (short *dst, const short *src, int n)
{
   /* the model blob: a multiple of 8 shorts, 16-byte aligned wherever it lives */
   const uint4 *s4 = (const uint4 *) src;
   uint4       *d4 = (uint4 *) dst;

   for (int i = threadIdx.x; i < n / 8; i += NT)
      d4 [i] = s4 [i];
}

/*****************************************************************************
	motion compensation of predicted frames  (codec/mwfa.c, codec/prediction.c)
*****************************************************************************/

/* MPEG's code lengths of the vector components (mwfa.c:40-52, second column) */
__constant__ unsigned char c_mv_code_length [33] =
{11, 11, 11, 11, 11, 11, 10, 10, 10, 8, 8, 8, 7, 5, 4, 3, 1, 3, 4, 5, 7, 8, 8, 8, 10, 10, 10,
 11, 11, 11, 11, 11, 11};

__device__ __forceinline__ float *
norms_of_level (const DevParams &P, const TileWs &W, int level, int backward = 0)
{
   return GP (W.norms) + ((size_t) backward * (P.p_max - P.p_min + 1) + (level - P.p_min))
			 * (4 * P.sr * P.sr);
}

/*
 *  fill_norms_table (mwfa.c:544-602) for the block (x0, y0) of 'level': for every vector of the
 *  search range the squared norm of (original - displaced reference) / 16 (get_mcpe / mcpe_norm,
 *  mwfa.c:604-684), each summed in row order in fp32 by one thread; 0 for vectors that leave the
 *  frame.
 */
template <int NT>
__device__ void
cta_fill_norms (const DevParams &P, const TileWs &W, unsigned x0, unsigned y0, int level)
{
   const int	  sr = P.sr, nd = 4 * sr * sr;
   const int	  bw = (int) width_of_level (level), bh = (int) height_of_level (level);
   const int16_t *orig = GP (W.pix);

   for (int dir = 0; dir < (P.motion == 2 ? 2 : 1); dir++)
   for (int index = threadIdx.x; index < nd; index += NT)
   {
      float	    *out  = norms_of_level (P, W, level, dir);
      const int16_t *past = dir ? GP (W.future) : GP (W.past);
      const int mx = index % (2 * sr) - sr, my = index / (2 * sr) - sr;
      float	norm = 0.0f;

      if ((int) x0 + mx >= 0 && (int) x0 + mx + bw <= P.width
	  && (int) y0 + my >= 0 && (int) y0 + my + bh <= P.height)
	 for (int y = 0; y < bh; y++)
	 {
	    const int16_t *o = orig + (size_t) (y0 + y) * P.width + x0;
	    const int16_t *r = past + (size_t) ((int) y0 + my + y) * P.width + ((int) x0 + mx);

	    for (int x = 0; x < bw; x++)
	    {
	       const int16_t d = (int16_t) (o [x] - r [x]);
	       const int     q = d / 16;

	       norm += (float) (q * q);
	    }
	 }
      out [index] = norm;
   }
   __syncthreads ();
}

/*
 *  find_best_mv (mwfa.c:686-795), full-pixel search: the first vector in index order with the
 *  least costs norm + (bits of the components) * price.  Result in h->best_i.
 */
template <int NT>
__device__ void
cta_find_best_mv (const DevParams &P, const TileWs &W, const Sh &sh, unsigned x0, unsigned y0,
		  int level, float price, int backward)
{
   const int	sr = P.sr, nd = 4 * sr * sr;
   const int	bw = (int) width_of_level (level), bh = (int) height_of_level (level);
   const float *norms = norms_of_level (P, W, level, backward);
   float	best  = FB_MAXCOSTS;
   int		besti = -1;

   for (int index = threadIdx.x; index < nd; index += NT)
   {
      const int mx = index % (2 * sr) - sr, my = index / (2 * sr) - sr;

      if ((int) x0 + mx >= 0 && (int) y0 + my >= 0
	  && (int) x0 + mx + bw <= P.width && (int) y0 + my + bh <= P.height)
      {
	 const float costs = norms [index] + ((float) c_mv_code_length [mx + sr]
					      + (float) c_mv_code_length [my + sr]) * price;
	 if (costs < best)
	 {
	    best  = costs;
	    besti = index;
	 }
      }
   }
   /* the least (costs, index) pair: inside the warps by shuffles, then over the warps' results
      (the pursuit's work arrays are idle here) */
#pragma unroll
   for (int o = 16; o > 0; o >>= 1)
   {
      const float c = __shfl_xor_sync (0xffffffffu, best, o);
      const int	  i = __shfl_xor_sync (0xffffffffu, besti, o);

      if (i >= 0 && (besti < 0 || c < best || (c == best && i < besti)))
      {
	 best  = c;
	 besti = i;
      }
   }
   if ((threadIdx.x & 31) == 0)
   {
      sh.num [threadIdx.x >> 5]		   = best;
      ((int *) sh.den) [threadIdx.x >> 5] = besti;
   }
   __syncthreads ();
   if (threadIdx.x == 0)
   {
      float m  = FB_MAXCOSTS;
      int   mi = -1;

      for (int t = 0; t < NT / 32; t++)
      {
	 const float c = sh.num [t];
	 const int   i = ((const int *) sh.den) [t];

	 if (i >= 0 && (c < m || (c == m && i < mi)))
	 {
	    m  = c;
	    mi = i;
	 }
      }
      sh.h->best_i = mi;
      sh.h->best_c = m;
   }
   __syncthreads ();
}

/*
 *  The prediction error of the block (x0, y0) of 'level' for the vector (mx, my) (get_mcpe,
 *  mwfa.c:604-649), cut to bintree order as floats (cut_to_bintree, subdivide.c:504-541), with
 *  the sums of squares of its nodes: node 0 is the whole block (top = level).
 */
template <int NT>
__device__ void
cta_mcpe_range (const DevParams &P, const TileWs &W, const Sh &cs, unsigned x0, unsigned y0,
		int level, int mctype, int mx, int my, int bx, int by)
{
   const int	  tid  = threadIdx.x;
   const unsigned size = 1u << level;
   const int16_t *orig = GP (W.pix), *past = GP (W.past), *fut = GP (W.future);

   for (unsigned i = tid; i < size; i += NT)
   {
      unsigned yy = 0, xx = 0;
      for (int b = 0; b < 11; b++)
      {
	 yy |= ((i >> (2 * b)) & 1u) << b;
	 xx |= ((i >> (2 * b + 1)) & 1u) << b;
      }
      const unsigned px = x0 + xx, py = y0 + yy;
      int ref;

      /* get_mcpe (mwfa.c:604-649): one reference block, or the mean of two (truncating) */
      if (mctype == 1)
	 ref = past [(size_t) ((int) py + my) * P.width + ((int) px + mx)];
      else if (mctype == 2)
	 ref = fut [(size_t) ((int) py + by) * P.width + ((int) px + bx)];
      else
	 ref = ((int) past [(size_t) ((int) py + my) * P.width + ((int) px + mx)]
		+ (int) fut [(size_t) ((int) py + by) * P.width + ((int) px + bx)]) / 2;
      const int16_t  d	= (int16_t) (orig [(size_t) py * P.width + px] - ref);
      cs.pixels [i] = (float) ((int) d / 16);
   }
   __syncthreads ();
   {
      const unsigned nleaf = 1u << (level - P.lmin);
      const unsigned len   = 1u << P.lmin;

      for (unsigned k = tid; k < nleaf; k += NT)
      {
	 int acc = 0;
	 for (unsigned i = 0; i < len; i++)
	 {
	    int v = (int) cs.pixels [k * len + i];
	    acc += v * v;
	 }
	 cs.norm_i [nleaf - 1 + k] = acc;
      }
      __syncthreads ();
      for (int l = P.lmin + 1; l <= level; l++)
      {
	 const unsigned nn    = 1u << (level - l);
	 const unsigned node0 = nn - 1;
	 for (unsigned k = tid; k < nn; k += NT)
	    cs.norm_i [node0 + k] = cs.norm_i [2 * (node0 + k) + 1] + cs.norm_i [2 * (node0 + k) + 2];
	 __syncthreads ();
      }
   }
}

/*
 *  subtract_mc (mwfa.c:156-299): for every motion compensated range of the luminance tree the
 *  displaced block of the reference frame(s) is taken off the Cb and Cr planes of the frame -- with
 *  the vector components rounded towards zero to even numbers, (v / 2) * 2, which restore_mc on the
 *  decoding side does not do; interpolated blocks as (a + b) / 2, 16-bit wrapping arithmetic.  The
 *  ranges are disjoint: no barrier between them.
 */
template <int NT>
__device__ void
cta_subtract_mc (const DevParams &P, const TileWs &W, unsigned states)
{
   const int	tid   = threadIdx.x;
   const size_t plane = (size_t) P.width * P.height;

   for (unsigned a = 2 * 3; a < 2 * states; a++)
   {
      const int type = GP (W.mv_type) [a];

      if (type == 0)
	 continue;
      const int level = (int) GP (W.level_of_state) [a >> 1] - 1;
      const int bw = (int) width_of_level (level), bh = (int) height_of_level (level);
      const int x0 = GP (W.x) [a], y0 = GP (W.y) [a];
      const int fx = (GP (W.mv_fx) [a] / 2) * 2, fy = (GP (W.mv_fy) [a] / 2) * 2;
      const int bx = (GP (W.mv_bx) [a] / 2) * 2, by = (GP (W.mv_by) [a] / 2) * 2;

      for (int i = tid; i < 2 * bw * bh; i += NT)
      {
	 const int	band = 1 + i / (bw * bh), k = i % (bw * bh);
	 const int	x = x0 + k % bw, y = y0 + k / bw;
	 int16_t       *o    = const_cast<int16_t *> (GP (W.pix)) + band * plane + (size_t) y * P.width + x;
	 const int16_t *past = GP (W.past) + band * plane, *fut = GP (W.future) + band * plane;
	 int		ref;

	 if (type == 1)
	    ref = past [(size_t) (y + fy) * P.width + x + fx];
	 else if (type == 2)
	    ref = fut [(size_t) (y + by) * P.width + x + bx];
	 else
	    ref = ((int) past [(size_t) (y + fy) * P.width + x + fx]
		   + (int) fut [(size_t) (y + by) * P.width + x + bx]) / 2;
	 *o = (int16_t) (*o - ref);
      }
   }
   __threadfence ();
   __syncthreads ();
}

/*
 *  nd_prediction (prediction.c:409-421): the difference between the range and its DC prediction
 *  becomes the pixel block of the nested pass.  src: the range inside the outer block (bintree
 *  order); dc = -weight * images_of_state [0][0].  The differences are no integers: the node norms
 *  of the nested pass are the reference's left-to-right fp32 sums, worked out here for all nodes at
 *  once.
 */
template <int NT>
__device__ void
cta_nd_range (const DevParams &P, const TileWs &W, const Sh &outer, const Sh &cs, unsigned address,
	      int level, float dc)
{
   const int	  tid  = threadIdx.x;
   const unsigned size = 1u << level;
   const float	 *src  = outer.pixels + ((size_t) address << level);

   for (unsigned i = tid; i < size; i += NT)
      cs.pixels [i] = src [i] + dc;
   __syncthreads ();
   /* one thread per node of the block's tree, each its own left-to-right sum (approx.c:388-389); the
      table holds the float's bits with the sign bit set -- an integer sum is never negative */
   for (unsigned k = tid; k < (2u << (level - P.lmin)) - 1; k += NT)
   {
      const unsigned d	  = 31u - (unsigned) __clz ((int) (k + 1));	/* depth of node k */
      const unsigned lvl  = (unsigned) level - d;
      const float   *px	  = cs.pixels + ((size_t) (k + 1 - (1u << d)) << lvl);
      float	     norm = 0;

      for (unsigned i = 0; i < (1u << lvl); i++)
	 norm += px [i] * px [i];
      cs.norm_i [k] = (int) (__float_as_uint (norm) | 0x80000000u);
   }
   __syncthreads ();
}

/*****************************************************************************
			  cluster per stream: jobs
*****************************************************************************/

/*
 *  Rank 0: hand the job in h->job (filled by thread 0) to the helper blocks, which wait at the
 *  cluster barrier.  with_models: the pursuits of a spine also need the probability models and
 *  the entries of the pool list written since the last time.  Everything is stored into the
 *  helpers' shared memory BEFORE the barrier, so nothing of rank 0 is read afterwards.
 */
template <int NT>
__device__ void
cta_cluster_post (const DevParams &P, const Sh &sh, bool with_models)
{
   ShHdr    *h	 = sh.h;
   const int tid = threadIdx.x;
   const int C	 = P.cluster;
   const int jw	 = (int) (sizeof (ClJob) / 4);

   LAP (h, LAP_CTRL);
   if (tid == 0)
      h->job.tswap = h->tswap;
   __syncthreads ();
   for (int it = tid; it < (C - 1) * jw; it += NT)
   {
      const unsigned r = 1u + (unsigned) (it / jw);
      const int	     i = it % jw;

      ((unsigned *) &cl_map (h, r)->job) [i] = ((const unsigned *) &h->job) [i];
   }
   if (with_models)
   {
      const int bw = P.blob_len / 8;
      const int lo = h->pool_lo, np = (int) h->job.pool_n - lo;

      for (int it = tid; it < (C - 1) * bw; it += NT)
      {
	 const unsigned r = 1u + (unsigned) (it / bw);
	 const int	i = it % bw;

	 ((uint4 *) cl_map (sh.blob, r)) [i] = ((const uint4 *) sh.blob) [i];
      }
      if (np > 0)
	 for (int it = tid; it < (C - 1) * np; it += NT)
	 {
	    const unsigned r = 1u + (unsigned) (it / np);
	    const int	   i = lo + it % np;

	    cl_map (sh.pool, r) [i] = sh.pool [i];
	 }
      __syncthreads ();
      if (tid == 0)
	 h->pool_lo = (int) h->job.pool_n;
   }
   cl_sync ();
   LAP (h, LAP_CLUSTER);
}

/* one pursuit of a spine on this block; the result goes to slot k of rank 0 */
template <int NT>
__device__ void
cta_spine_pursuit (const DevParams &P, const TileWs &W, const Sh &sh, int k, SpecRes *dst)
{
   ShHdr	  *h  = sh.h;
   const SpineNode nd = h->job.node [k];

   cta_matching_pursuit<NT> (P, W, sh, h->mp, nd.level, nd.image, nd.norm, nd.tree_bits, nd.mv_tree_bits,
			     h->job.price, FB_RANGE, -1);
   for (int i = threadIdx.x; i < (int) (sizeof (MpRes) / 4); i += NT)
      ((unsigned *) &dst->mp) [i] = ((const unsigned *) &h->mp) [i];
   if (threadIdx.x == 0)
   {
      dst->steps = h->w.st_steps;
      dst->pass2 = h->w.st_pass2;
      dst->D	 = (unsigned) h->w.D;
   }
   __syncthreads ();
}

/* ranks > 0: serve rank 0 until it says the frame is done */
template <int NT, bool MOTION>
__device__ void
cta_helper_loop (const DevParams &P, TileWs &W, const Sh &sh, unsigned rank)
{
   ShHdr    *h = sh.h;
   const int C = P.cluster;
   Sh	     shn = sh;		/* the nested pass of a predicted frame: delta models, prediction-error pixels */

   if (MOTION)
   {
      shn.blob	 = sh.blob + P.blob_half;
      shn.pixels = GP (W.pix2);
      shn.norm_i = GP (W.norm2);
      if (threadIdx.x == 0)
	 h->tswap = 0;
   }
   for (;;)
   {
      cl_sync ();			/* a job has been posted */
      const int	 type	= h->job.type;
      const bool nested = MOTION && h->job.nested;

      if (type == CJ_EXIT)
	 break;
      if (threadIdx.x == 0)
      {
	 h->states = h->job.states;
	 if (MOTION && h->job.tswap != h->tswap)
	 {
	    /* rank 0 works on the other product table (prediction.c:318-341): so do we */
	    float *t = W.T;

	    W.T	     = W.T2;
	    W.T2     = t;
	    h->tswap = h->job.tswap;
	 }
      }
      __syncthreads ();
      Sh cs = sh;

      if (nested)
      {
	 cs.blob   = shn.blob;
	 cs.pixels = shn.pixels;
	 cs.norm_i = shn.norm_i;
      }
      if (type == CJ_TINIT)
	 /* this block's share of the products of a new lc_max block; ends with the cluster barrier
	    after the last level */
	 cta_init_range<NT, true> (P, W, cs, (unsigned) h->job.x, (unsigned) h->job.y, h->job.band,
				   rank * NT, (unsigned) C * NT);
      else if (MOTION && type == CJ_TERR)
	 cta_compute_T<NT, true> (P, W, cs, 0, 0, h->job.level, h->job.level, rank * NT, (unsigned) C * NT);
      else if (type == CJ_APPEND)
      {
	 cta_state_products<NT> (P, W, sh, h->job.s, (int) rank, C);
	 cl_sync ();
      }
      else
      {
	 for (int k = (int) rank; k < h->job.n; k += C)
	    cta_spine_pursuit<NT> (P, W, cs, k, &cl_map (h, 0)->spec [k]);
	 cl_sync ();
      }
   }
}

/*****************************************************************************
		   the bintree recursion  (codec/subdivide.c:60-502)
*****************************************************************************/

enum { ST_CHILD_T = 16, ST_CHILD2 = 17,
       /* predicted frames */
       ST_AFTER_CHILD2 = 18, ST_NORMS_UP = 19, ST_FILL_CHILD = 20, ST_PRED_DONE = 21 };

/* where the activation record at 'depth' leaves its range: the parent's child slot, the root
   range, or -- for the root of a nested pass -- the prediction range of the record below */
/* (the records are addressed through the Sh copy in registers where one is at hand: the pointers
   in the header cost a shared-memory load each on thread 0's path) */
template <bool MOTION>
__device__ __forceinline__ RangeRes *
res_slot (const Sh &sh, int depth)
{
   if (depth == 0)
      return &sh.h->root;
   if (MOTION && depth == sh.h->nest_base)
      return &sh.fx [depth - 1].prange;
   return &sh.frames [depth - 1].child [sh.frames [depth - 1].label];
}

template <bool MOTION>
__device__ __forceinline__ RangeRes *
res_slot (ShHdr *h, int depth)
{
   if (depth == 0)
      return &h->root;
   if (MOTION && depth == h->nest_base)
      return &h->fx [depth - 1].prange;
   return &h->frames [depth - 1].child [h->frames [depth - 1].label];
}

__device__ __forceinline__ RangeX *
resx_slot (ShHdr *h, int depth)
{
   if (depth == 0)
      return &h->root_x;
   if (depth == h->nest_base)
      return &h->fx [depth - 1].prange_x;
   return &h->fx [depth - 1].child [h->frames [depth - 1].label];
}

/*
 *  Thread 0: run the scalar part of subdivide()'s control flow -- returns, cost
 *  bookkeeping between the two children, tree-model updates, child geometry, early
 *  exits -- until the block as a whole is needed again:
 *    ST_ENTER    a visible range of level >= 3: snapshot, linear combination, ...
 *    ST_CHILD_T  products of the states born in child 0 for child 1's subtree
 *    ST_DECIDE   restore / adopt models or append a new state (predicted frames: the
 *		  motion compensated alternative first)
 *    ST_NORMS_UP / ST_FILL_CHILD / ST_PRED_DONE  (predicted frames) norms tables, end of the
 *		  nested pass
 *    ST_DONE
 */
/*
 *  Thread 0: a range whose pursuit ran with the spine it belongs to (F.spec_k > 0) is entered
 *  without the block: what ST_ENTER does for it -- the range's bookkeeping, approximate_range
 *  after the pursuit (t0_range_epilogue), the start of the subdivision alternative -- is scalar
 *  work on the result in h->spec; the snapshots of the models were taken for the whole spine
 *  when it was started.  Leaves the range in ST_CHILD (first child next) or, at the lowest
 *  level, in ST_RETURN.
 */
template <bool MOTION>
__device__ void
t0_enter_speculated (const DevParams &P, const TileWs &W, const Sh &sh, Frame &F, FrameX *X, RangeRes *res,
		     int &state)
{
   ShHdr	 *h	= sh.h;
   const int	  level = F.level;
   const int	  k	= F.spec_k;
   const SpecRes &r	= h->spec [k];
   const bool	  leaf	= !MOTION && level <= h->lc_min;	/* (predicted frames: a third alternative follows) */
   RangeRes	 &lr	= leaf ? *res : F.lrange;
   float	  mvt	= 0.0f;

   if (MOTION)
   {
      /* what ST_ENTER sets up for the motion compensated alternative (subdivide.c:141-147) */
      mvt	   = h->job.node [k].mv_tree_bits;
      X->try_mc	   = mvt != 0.0f;
      /* (an intra frame with nondeterministic prediction: the prediction tree model has not changed
	 since the spine was started -- no child has returned -- so the bits of "not predicted here"
	 are those of this moment, as in ST_ENTER) */
      X->try_nd	   = P.motion == 3 && X->prediction && level >= P.p_min && level <= P.p_max;
      X->pred_done = 0;
      X->lrange.mv_tree_bits  = mvt;
      if (X->try_nd)
	 mvt = t0_ptree_bits (h, 1, level);
      X->lrange.mv_coord_bits = 0;
      X->lrange.mv_type = X->lrange.mv_fx = X->lrange.mv_fy = X->lrange.prediction = 0;
      X->lrange.mv_bx = X->lrange.mv_by = 0;
      X->r_mvt = mvt;
      X->r_mvc = 0;
      for (int label = 0; label < 2; label++)
      {
	 X->child [label].mv_tree_bits = X->child [label].mv_coord_bits = 0;
	 X->child [label].mv_type = X->child [label].mv_fx = X->child [label].mv_fy = 0;
	 X->child [label].mv_bx = X->child [label].mv_by = 0;
	 X->child [label].prediction = 0;
      }
   }

   F.states_snap     = h->states;
   if (MOTION)
      X->off_snap = h->hole_off;
   F.new_y_state [0] = F.new_y_state [1] = FB_RANGE;	/* (speculated ranges have no y-state) */
   lr.tree	   = FB_RANGE;
   lr.x		   = (unsigned short) F.x;
   lr.y		   = (unsigned short) F.y;
   lr.tree_bits	   = h->job.node [k].tree_bits;
   lr.matrix_bits  = 0;
   lr.weights_bits = 0;
   lr.err	   = 0;
   lr.into [0]	   = FB_NO_EDGE;
   h->mp = r.mp;
   h->mp_calls++;
   h->mp_steps += r.steps;
   h->pass2    += r.pass2;
   h->mp_bytes += 8ull * r.D + 4ull * r.D * r.steps;
   F.lincomb_costs = t0_range_epilogue (P, W, sh, h->mp, -1, F.max_costs, h->price, F.y_state, &lr, level,
					F.image, F.address, F.x, F.y, leaf ? (Frame *) 0 : &F);
   if (leaf)
   {
      state = ST_RETURN;
      return;
   }
   if (level <= h->lc_min)	/* (predicted frames) no subdivision: on to the motion compensated alternative */
   {
      F.subdivide_costs = FB_MAXCOSTS;
      state		= ST_DECIDE;
      return;
   }
   /* alternative 2: recursive subdivision (subdivide.c:243-272) */
   F.r_tree_bits     = h->spec_rbits [k];
   F.r_matrix_bits   = 0;
   F.r_weights_bits  = 0;
   F.r_err	     = 0;
   F.subdivide_costs = (F.r_tree_bits + F.r_weights_bits + F.r_matrix_bits + mvt + 0.0f + 0.0f + 0.0f)
		       * h->price;
   F.label = 0;
   for (int label = 0; label < 2; label++)
   {
      RangeRes &c = F.child [label];

      c.tree = 0;
      c.into [0] = 0;
      c.err = c.tree_bits = c.matrix_bits = c.weights_bits = 0;
   }
   state = ST_CHILD;
}

template <bool MOTION, bool CLU>
__device__ void
t0_advance (const DevParams &P, const TileWs &W, const Sh &sh, int &state, int &depth)
{
   ShHdr *h = sh.h;

   for (;;)
   {
      Frame &F = sh.frames [depth];

      if (state == ST_ENTER)
      {
	 RangeRes *res = res_slot<MOTION> (sh, depth);

	 res->into [0] = FB_NO_EDGE;
	 res->tree     = FB_RANGE;
	 if (F.level < 3)			/* subdivide.c:113 */
	 {
	    h->ret_costs = FB_MAXCOSTS;
	    state	 = ST_RETURN;
	 }
	 else if (F.x >= (unsigned) P.width || F.y >= (unsigned) P.height)
	 {
	    h->ret_costs = 0;			/* subdivide.c:133-135 */
	    state	 = ST_RETURN;
	 }
	 else if (CLU && F.spec_k > 0)
	 {
	    /* (the nested pass of a predicted frame: the delta models) */
	    Sh cs = sh;

	    if (MOTION && h->nest_base >= 0 && depth >= h->nest_base)
	       cs.blob = sh.blob + P.blob_half;
	    t0_enter_speculated<MOTION> (P, W, cs, F, MOTION ? &sh.fx [depth] : (FrameX *) 0, res, state);
	 }
	 else
	    return;
      }
      else if (state == ST_RETURN)
      {
	 if (depth == 0)
	 {
	    state = ST_DONE;
	    return;
	 }
	 if (MOTION && depth == h->nest_base)
	 {
	    depth--;				/* back in mc_prediction (prediction.c:318) */
	    state = ST_PRED_DONE;
	    return;
	 }
	 sh.frames [depth - 1].subdivide_costs += h->ret_costs;
	 depth--;
	 state = ST_AFTER_CHILD;
      }
      else if (state == ST_AFTER_CHILD)
      {
	 /* update_norms_table (prediction.c:213-238) */
	 if (MOTION && sh.fx [depth].try_mc && F.level > P.p_min)
	 {
	    state = ST_NORMS_UP;
	    return;
	 }
	 state = ST_AFTER_CHILD2;
      }
      else if (state == ST_AFTER_CHILD2)
      {
	 const int label = F.label;
	 RangeRes &c	 = F.child [label];

	 /* progress meter (subdivide.c:323-337): (global_address + 1) * 100.0 / 2^k, truncated --
	    exact in integers.  Nothing is printed here; the host replays the values. */
	 {
	    /* (32 bits are enough: at most 2^22 ranges, times 100) */
	    const unsigned pos = F.gaddr * 2 + (unsigned) label + 1;
	    const unsigned np  = (pos * 100u) >> (P.level - (F.level - 1));

	    if (np > h->percent)
	    {
	       h->percent = np;
	       h->progress [(np >> 5) & 3] |= 1u << (np & 31);
	    }
	 }

	 if (F.subdivide_costs >= fmin2 (F.lincomb_costs, F.max_costs))
	 {
	    F.subdivide_costs = FB_MAXCOSTS;	/* subdivide.c:355-359 */
	    state	      = ST_DECIDE;
	    return;
	 }
	 F.r_err	  += c.err;
	 F.r_tree_bits	  += c.tree_bits;
	 F.r_matrix_bits  += c.matrix_bits;
	 F.r_weights_bits += c.weights_bits;
	 if (MOTION)
	 {
	    sh.fx [depth].r_mvt += sh.fx [depth].child [label].mv_tree_bits;
	    sh.fx [depth].r_mvc += sh.fx [depth].child [label].mv_coord_bits;
	 }
	 /* tree_update (bintree.c:35-53) of the tree model and of the prediction tree model
	    (subdivide.c:370-373), which only nondeterministic prediction reads */
	 if (c.tree != FB_RANGE)
	    h->tree_counts [F.level - 1]++;
	 h->tree_total [F.level - 1]++;
	 if (MOTION)
	 {
	    if (!sh.fx [depth].child [label].prediction)
	       h->tree_counts [F.level - 1] += 0x10000u;
	    h->tree_total [F.level - 1] += 0x10000u;
	 }
	 F.label = label + 1;
	 if (F.label >= 2)
	 {
	    state = ST_DECIDE;
	    return;
	 }
	 state = ST_CHILD;
      }
      else if (state == ST_CHILD || state == ST_CHILD2)
      {
	 const int	label = F.label;
	 const int	level = F.level;
	 const unsigned cimg  = F.image * 2 + label + 1;
	 const unsigned cadr  = F.address * 2 + label;
	 const unsigned cx    = (level & 1) ? F.x : F.x + label * width_of_level (level - 1);
	 const unsigned cy    = (level & 1) ? F.y + label * height_of_level (level - 1) : F.y;

	 /* products of the states born in child 0 (subdivide.c:295-297) */
	 if (state == ST_CHILD && label && level <= P.lc_max && F.states_snap < h->states)
	 {
	    state = ST_CHILD_T;
	    return;
	 }
	 const float remaining = fmin2 (F.lincomb_costs, F.max_costs) - F.subdivide_costs;

	 F.child [label].x = (unsigned short) cx;
	 F.child [label].y = (unsigned short) cy;
	 if (remaining > 0)
	 {
	    Frame &C = sh.frames [depth + 1];

	    C.max_costs = remaining;
	    C.x		= cx;
	    C.y		= cy;
	    C.image	= cimg;
	    C.address	= cadr;
	    C.gaddr	= F.gaddr * 2 + label;
	    C.level	= level - 1;
	    C.y_state	= F.new_y_state [label];
	    /* the label-0 descendants of the head of a spine find their pursuit done */
	    C.spec_k	= (label == 0 && F.spec_k >= 0 && F.spec_k + 1 < h->spec_len) ? F.spec_k + 1 : -1;
	    if (MOTION)
	    {
	       sh.fx [depth + 1].delta	    = sh.fx [depth].delta;
	       sh.fx [depth + 1].prediction = sh.fx [depth].prediction;
	    }
	    depth++;
	    state = ST_ENTER;
	 }
	 else if (MOTION && sh.fx [depth].try_mc && level - 1 >= P.p_min)
	 {
	    state = ST_FILL_CHILD;	/* subdivide.c:331-333: the child's norms are still needed */
	    return;
	 }
	 else
	    state = ST_AFTER_CHILD;	/* subdivide() not called: costs unchanged */
      }
      else
	 return;
   }
}

template <int NT, bool MOTION>
__device__ void
cta_subdivide_band (const DevParams &P, TileWs &W, const Sh &sh, int band,
		    int root_y_state)
{
   ShHdr    *h	 = sh.h;
   const int tid = threadIdx.x;
   int	     it	 = 0;
   const bool CL = clustered_shape<NT, MOTION> () && P.cluster > 1;	/* helper blocks at hand */
   /* The batch shapes update the models the way the reference does: at once, the result kept aside
      as "lc" models until the range is decided (subdivide.c:226-237).  The clustered shape leaves
      the update to ST_DECIDE (t0_range_epilogue): its speculated ranges are entered by thread 0
      alone, which cannot copy models.  Measured on a full batch: 786 against 762 Mpx/s for the
      late update, whose serial part sits behind a read of the snapshot from global memory. */
   constexpr bool EAGER = !clustered_shape<NT, MOTION> ();
   const int SN	 = MOTION ? 3 : 2;	/* model snapshots per activation record */
   const int TS	 = MOTION ? 2 : 1;	/* tree-model snapshots per record */
   Sh	     shn = sh;			/* buffers and models of the nested (prediction error) pass */

   if (MOTION)
   {
      shn.blob	 = sh.blob + P.blob_half;
      shn.pixels = GP (W.pix2);
      shn.norm_i = GP (W.norm2);
   }
   if (tid == 0)
   {
      Frame &F = h->frames [0];
      int    st = ST_ENTER, dp = 0;

      F.max_costs = FB_MAXCOSTS;
      F.x = F.y = F.image = F.address = F.gaddr = 0;
      F.level	= P.level;
      F.spec_k	= -1;
      h->spec_len = 0;
      h->percent = 0;
      h->progress [0] = h->progress [1] = h->progress [2] = h->progress [3] = 0;
      F.y_state = root_y_state;
      h->band	= band;
      h->price	= band ? P.price * P.chroma_decrease : P.price;
      h->nest_base = -1;
      h->top	   = P.lc_max;
      if (MOTION)
      {
	 h->fx [0].delta      = 0;
	 h->fx [0].prediction = band == 0 && P.motion != FB200_FRAME_INTRA;	/* coder.c:806-807: the chroma bands are not predicted */
      }
      t0_advance<MOTION, clustered_shape<NT, MOTION> ()> (P, W, sh, st, dp);
      h->state [0]  = st;
      h->depthv [0] = dp;
   }
   __syncthreads ();

   for (;; it++)
   {
      const int state = h->state [it & 1];
      const int depth = h->depthv [it & 1];
      Frame    &F     = h->frames [depth];
      RangeRes *res   = res_slot<MOTION> (h, depth);
      int	nstate = state;		/* thread 0: state after this action */
      int	ndepth = depth;
      /* nested pass of a prediction: the delta models, the prediction error block */
      const bool nested = MOTION && h->nest_base >= 0 && depth >= h->nest_base;
      /* (a copy with the three members that differ selected one by one: a reference to one of two
	 structs would put both into local memory) */
      Sh	 cs	= sh;

      if (MOTION && nested)
      {
	 cs.blob   = shn.blob;
	 cs.pixels = shn.pixels;
	 cs.norm_i = shn.norm_i;
      }

      if (state == ST_DONE || h->status != FB200_OK)
	 break;
      LAP (h, LAP_CTRL);

      /* range x state products: the whole block when an lc_max range is entered
	 (init_range, subdivide.c:612-644), or the states born in child 0 for child 1's
	 subtree (subdivide.c:295-297); one call site = one copy of the code */
      if (state == ST_CHILD_T || (state == ST_ENTER && F.level == P.lc_max))
      {
	 const long long t0c   = clock64 ();
	 const bool	 block = state == ST_ENTER;

	 if (block && CL)
	 {
	    /* every block of the cluster takes its share of the states */
	    if (tid == 0)
	    {
	       h->job.type   = CJ_TINIT;
	       h->job.nested = nested;
	       h->job.states = h->states;
	       h->job.x	     = (int) F.x;
	       h->job.y	     = (int) F.y;
	       h->job.band   = band;
	    }
	    cta_cluster_post<NT> (P, sh, false);
	    cta_init_range<NT, clustered_shape<NT, MOTION> ()> (P, W, cs, F.x, F.y, band, 0, (unsigned) P.cluster * NT);
	 }
	 else if (block)
	    cta_init_range<NT> (P, W, cs, F.x, F.y, band);
	 else
	    cta_compute_T<NT, clustered_shape<NT, MOTION> ()> (P, W, cs, F.states_snap, F.image * 2 + F.label + 1, F.level - 1,
			       MOTION ? h->top : P.lc_max);
	 if (tid == 0)
	 {
	    if (block)
	       F.address = F.image = 0;
	    h->cyc_T += clock64 () - t0c;
	 }
      }

      switch (state)
      {
	 case ST_ENTER:
	 {
	    const int level = F.level;
	    short    *snap  = (sh.snaps ? sh.snaps : GP (W.snap)) + (size_t) depth * SN * P.blob_len;
	    unsigned *tsnap = sh.tsnap + (size_t) depth * TS * 2 * FB200_MAXLEVEL;
	    /*
	     *  A range of the lowest level cannot be subdivided: its linear combination is the
	     *  result or the range fails.  An accepted approximation leaves the models it has
	     *  updated, a rejected one leaves them untouched (approx.c:219-252), and nothing else
	     *  has changed -- exactly what restoring / adopting the snapshots of
	     *  subdivide.c:409-467 amounts to -- so no snapshot is taken and the range returns
	     *  from here.  (Predicted frames keep the general path: a third alternative follows.)
	     */
#ifdef FB200_X_NOLEAF
	    const bool leaf = false;
#else
	    const bool leaf = !MOTION && level <= h->lc_min && level <= P.lc_max;
#endif

	    /* snapshot of the models (subdivide.c:188-194); tree_counts and tree_total are
	       adjacent in the header */
	    if (!leaf)
	       cta_copy_s16<NT> (snap, sh.blob, P.blob_len);
	    if (!leaf && tid < 32)
	    {
	       const unsigned *tm = (const unsigned *) ((const char *) h + offsetof (ShHdr, tree_counts));

	       for (int i = tid; i < 2 * FB200_MAXLEVEL; i += 32)
		  tsnap [i] = tm [i];
	       __syncwarp ();
	    }
	    if (tid == 0)
	    {
	       float mvt = 0.0f;

	       if (MOTION)
	       {
		  FrameX &X = h->fx [depth];

		  /* motion compensation allowed for this range? (subdivide.c:141-147) */
		  X.try_mc = P.motion != 3 && X.prediction && level >= P.p_min && level <= P.p_max
			     && F.x + width_of_level (level) <= (unsigned) P.width
			     && F.y + height_of_level (level) <= (unsigned) P.height;
		  /* an intra frame with nondeterministic prediction (subdivide.c:149-151).  Its bits
		     travel in the motion fields, which are zero in such a frame: nd_tree_bits in
		     mv_tree_bits, nd_weights_bits in mv_coord_bits (the sums of cwfa.h's seven bit
		     counts are the same numbers: adding 0 is exact) */
		  X.try_nd = P.motion == 3 && X.prediction && level >= P.p_min && level <= P.p_max;
		  X.pred_done = 0;
		  mvt	      = X.try_mc ? 1.0f : 0.0f;	/* mc allowed but not used */
		  X.lrange.mv_tree_bits	 = mvt;
		  if (X.try_nd)
		     mvt = t0_ptree_bits (h, 1, level);	/* rrange.nd_tree_bits (subdivide.c:259-260) */
		  X.lrange.mv_coord_bits = 0;
		  X.lrange.mv_type = X.lrange.mv_fx = X.lrange.mv_fy = X.lrange.prediction = 0;
		  X.lrange.mv_bx = X.lrange.mv_by = 0;
		  X.r_mvt = mvt;
		  X.r_mvc = 0;
		  for (int label = 0; label < 2; label++)
		  {
		     X.child [label].mv_tree_bits = X.child [label].mv_coord_bits = 0;
		     X.child [label].mv_type = X.child [label].mv_fx = X.child [label].mv_fy = 0;
		     X.child [label].mv_bx = X.child [label].mv_by = 0;
		     X.child [label].prediction = 0;
		  }
	       }
	       F.states_snap = h->states;
	       if (MOTION)
		  h->fx [depth].off_snap = h->hole_off;
	       /* y states of the children (subdivide.c:172-183) */
	       if (!leaf)
		  for (int label = 0; label < 2; label++)
		     F.new_y_state [label] = (band != 0 && F.y_state != FB_RANGE)
					     ? (int) GP (W.tree) [2 * F.y_state + label] : FB_RANGE;
	       F.lincomb_costs = FB_MAXCOSTS;
	       if (level <= P.lc_max)
	       {
		  RangeRes &lr = leaf ? *res : F.lrange;	/* a leaf fills its result slot directly */

		  lr.tree	  = FB_RANGE;
		  lr.x		  = (unsigned short) F.x;
		  lr.y		  = (unsigned short) F.y;
		  lr.tree_bits	  = t0_tree_bits (h, 0, level);
		  lr.matrix_bits  = 0;
		  lr.weights_bits = 0;
		  lr.err	  = 0;
		  lr.into [0]	  = FB_NO_EDGE;
	       }
	    }
	    else if (clustered_shape<NT, MOTION> () && tid == 32 && !leaf && level > h->lc_min)
	       /* bits of the "subdivided" symbol (subdivide.c:243-248), next to thread 0's: the tree
		  model does not change before they are used */
	       F.r_tree_bits = t0_tree_bits (h, 1, level);
	    __syncthreads ();
	    /* clear_norms_table (prediction.c:195-211) */
	    if (MOTION && h->fx [depth].try_mc && level > P.p_min)
	    {
	       for (int dir = 0; dir < (P.motion == 2 ? 2 : 1); dir++)
	       {
		  float *nt = norms_of_level (P, W, level, dir);

		  for (int i = tid; i < 4 * P.sr * P.sr; i += NT)
		     nt [i] = 0.0f;
	       }
	    }
	    /* alternative 1: linear combination (subdivide.c:200-221) */
	    if (level <= P.lc_max)
	    {
	       const long long t0c   = clock64 ();
	       const int       sk    = CL ? F.spec_k : 0;
	       bool	       spine = false;

	       if (CL && sk < 0 && F.y_state < 0 && !P.second_domain_block)
	       {
		  /*
		   *  Head of a spine: the ranges reached from here by label 0 alone, down to
		   *  lc_min_level, are approximated with the models and states of this moment
		   *  (subdivide.c:188-237: the models are put back before the first child is
		   *  entered; states are appended only when a child returns) -- their pursuits run
		   *  now, one per block of the cluster, next to this range's own.
		   */
		  int n = level - h->lc_min + 1;

		  if (n > FB_SPINE_MAX)
		     n = FB_SPINE_MAX;
		  if (n >= 2)
		  {
		     if (tid == 0)
		     {
			h->job.type   = CJ_SPINE;
			h->job.nested = nested;
			h->job.n      = n;
			h->job.states = h->states;
			h->job.pool_n = BLOB_U16 (sh, MB_N);
			h->job.price  = h->price;
			h->spec_len   = n;
		     }
		     if (tid < n)	/* one thread per node: the log2 of the tree bits side by side */
		     {
			const int  k  = tid;
			SpineNode &nd = h->job.node [k];

			nd.level     = level - k;
			nd.image     = ((F.image + 1) << k) - 1;
			nd.address   = F.address << k;
			nd.tree_bits = k ? t0_tree_bits (h, 0, level - k) : F.lrange.tree_bits;
			nd.norm	     = t0_node_norm (P, cs, nd.image, nd.address, nd.level);
			/* motion compensation allowed for the range? (subdivide.c:141-147; the range shares
			   its corner with this one, so it lies inside the frame if this one does) */
			nd.mv_tree_bits = MOTION && P.motion != 3 && h->fx [depth].prediction && level - k >= P.p_min
					  && level - k <= P.p_max
					  && F.x + width_of_level (level - k) <= (unsigned) P.width
					  && F.y + height_of_level (level - k) <= (unsigned) P.height ? 1.0f : 0.0f;
		     }
		     else if (tid >= 32 && tid < 32 + n)
			h->spec_rbits [tid - 32] = t0_tree_bits (h, 1, level - (tid - 32));
		     /* the snapshots of the models (subdivide.c:188-194) for the ranges below: they are
			entered with the models and the tree model of this moment (t0_enter_speculated) */
		     for (int i = tid; i < (n - 1) * (P.blob_len / 8); i += NT)
		     {
			const int k = 1 + i / (P.blob_len / 8), j = i % (P.blob_len / 8);

			((uint4 *) (snap + (size_t) k * SN * P.blob_len)) [j] = ((const uint4 *) sh.blob) [j];
		     }
		     for (int i = tid; i < (n - 1) * 2 * FB200_MAXLEVEL; i += NT)
		     {
			const int	k  = 1 + i / (2 * FB200_MAXLEVEL), j = i % (2 * FB200_MAXLEVEL);
			const unsigned *tm = (const unsigned *) ((const char *) h + offsetof (ShHdr, tree_counts));

			(tsnap + (size_t) k * TS * 2 * FB200_MAXLEVEL) [j] = tm [j];
		     }
		     /* clear_norms_table (prediction.c:195-211) of the ranges below, which are entered
			without the block: nothing touches those tables before */
		     if (MOTION && h->fx [depth].prediction)
			for (int k = 1; k < n; k++)
			{
			   const int lk = level - k;

			   if (lk > P.p_min && lk <= P.p_max
			       && F.x + width_of_level (lk) <= (unsigned) P.width
			       && F.y + height_of_level (lk) <= (unsigned) P.height)
			      for (int dir = 0; dir < (P.motion == 2 ? 2 : 1); dir++)
			      {
				 float *nt = norms_of_level (P, W, lk, dir);

				 for (int i = tid; i < 4 * P.sr * P.sr; i += NT)
				    nt [i] = 0.0f;
			      }
			}
		     cta_cluster_post<NT> (P, sh, true);
		     if (tid == 0)
			F.spec_k = 0;	/* (after the barriers of the post: every thread has read it) */
		     /* a spine longer than the cluster: this block's further nodes */
		     for (int k = P.cluster; k < n; k += P.cluster)
			cta_spine_pursuit<NT> (P, W, cs, k, &h->spec [k]);
		     spine = true;
		  }
	       }
	       cta_approximate_range<NT> (P, W, cs, F.max_costs, h->price, F.y_state,
					  leaf ? res : &F.lrange, level, F.image, F.address, F.x, F.y,
					  MOTION ? h->fx [depth].lrange.mv_tree_bits : 0.0f,
					  leaf || EAGER ? (Frame *) 0 : &F, sk > 0 ? sk : 0);
	       if (spine)
	       {
		  cl_sync ();		/* the helpers' results are in h->spec */
		  LAP (h, LAP_CLUSTER);
	       }
	       if (tid == 0)
	       {
		  F.lincomb_costs = h->ret_costs;
		  h->cyc_mp += clock64 () - t0c;
	       }
	    }
	    if (leaf)
	    {
	       /* the result slot holds the range (or nothing), ret_costs its costs (or MAXCOSTS) */
	       if (tid == 0)
		  nstate = ST_RETURN;
	       break;
	    }
	    /* (the models have not taken the linear combination: no "lc" models to keep, nothing
	       to restore, subdivide.c:226-237 -- see t0_range_epilogue) */
	    if (EAGER)
	       for (int i = tid; i < P.blob_len / 8; i += NT)
	       {
		  const uint4 lc = ((const uint4 *) sh.blob) [i];
		  const uint4 sn = ((const uint4 *) snap) [i];

		  ((uint4 *) (snap + P.blob_len)) [i] = lc;
		  ((uint4 *) sh.blob) [i]	      = sn;
	       }
	    if (tid == 0)
	    {
	       if (level > h->lc_min)
	       {
		  /* alternative 2: recursive subdivision (subdivide.c:243-272); the shape with an SM
		     to itself had thread 32 work out r_tree_bits next to thread 0's bits */
		  if (!clustered_shape<NT, MOTION> ())
		     F.r_tree_bits = t0_tree_bits (h, 1, level);
		  F.r_matrix_bits  = 0;
		  F.r_weights_bits = 0;
		  F.r_err	   = 0;
		  F.subdivide_costs = (F.r_tree_bits + F.r_weights_bits + F.r_matrix_bits
				       + (MOTION ? h->fx [depth].r_mvt : 0.0f) + 0.0f + 0.0f + 0.0f)
				      * h->price;
		  F.label = 0;
		  for (int label = 0; label < 2; label++)
		  {
		     RangeRes &c = F.child [label];
		     c.tree = 0;	/* memset (child, 0) */
		     c.into [0] = 0;
		     c.err = c.tree_bits = c.matrix_bits = c.weights_bits = 0;
		  }
		  nstate = ST_CHILD;
	       }
	       else
	       {
		  F.subdivide_costs = FB_MAXCOSTS;
		  nstate	    = ST_DECIDE;
	       }
	    }
	    break;
	 }

	 case ST_CHILD_T:
	 {
	    if (tid == 0)
	       nstate = ST_CHILD2;
	    break;
	 }

	 case ST_NORMS_UP:		/* update_norms_table (prediction.c:213-238) */
	 {
	    if (MOTION)
	    {
	       for (int dir = 0; dir < (P.motion == 2 ? 2 : 1); dir++)
	       {
		  float	      *up = norms_of_level (P, W, F.level, dir);
		  const float *lo = norms_of_level (P, W, F.level - 1, dir);

		  for (int i = tid; i < 4 * P.sr * P.sr; i += NT)
		     up [i] += lo [i];
	       }
	       if (tid == 0)
		  nstate = ST_AFTER_CHILD2;
	    }
	    break;
	 }

	 case ST_FILL_CHILD:		/* fill_norms_table for a child that is not visited */
	 {
	    if (MOTION)
	    {
	       cta_fill_norms<NT> (P, W, F.child [F.label].x, F.child [F.label].y, F.level - 1);
	       if (tid == 0)
		  nstate = ST_AFTER_CHILD;
	    }
	    break;
	 }

	 case ST_PRED_DONE:		/* mc_prediction after the nested subdivide (prediction.c:318-360) */
	 {
	    if (MOTION)
	    {
	       FrameX	  &X	 = h->fx [depth];
	       short	  *snap	 = (sh.snaps ? sh.snaps : GP (W.snap)) + (size_t) depth * SN * P.blob_len;
	       unsigned	  *tsnap = sh.tsnap + (size_t) depth * TS * 2 * FB200_MAXLEVEL;
	       const float costs = X.pcosts + h->ret_costs;
	       /* (nondeterministic prediction is only taken with a subdivided difference,
		  prediction.c:447) */
	       const bool  win	 = costs < X.max_pred && !(X.try_nd && X.prange.tree == FB_RANGE);

	       __syncthreads ();
	       if (tid == 0)
	       {
		  if (h->nest_base >= 0)	/* the outer tables come back */
		  {
		     float *t = W.T;
		     W.T      = W.T2;
		     W.T2     = t;
		     h->tswap ^= 1;
		  }
		  h->nest_base = -1;
		  h->top       = P.lc_max;
	       }
	       __syncthreads ();
	       if (win)
	       {
		  /* the products of the states born in the nested pass start from zero in the
		     outer tables (prediction.c:337-341) */
		  const unsigned first = X.last_state + 1, nnew = h->states - first;

		  for (unsigned i = tid; i < (unsigned) P.tn * nnew; i += NT)
		     GP (W.T) [(size_t) (i / nnew) * P.s_cap + first + i % nnew] = 0.0f;
		  /* the states of the split alternative stay behind as holes:
		     inert, never referenced; the host closes them */
		  for (unsigned s = F.states_snap + tid; s < X.rec_states; s += NT)
		  {
		     for (int label = 0; label < 2; label++)
		     {
			GP (W.into) [(size_t) (2 * s + label) * 6] = FB_NO_EDGE;
			GP (W.tree) [2 * s + label]		   = FB_RANGE;
			GP (W.mv_type) [2 * s + label]		   = 0;
		     }
		     GP (W.level_of_state) [s] = 255;
		  }
		  if (tid == 0)
		  {
		     RangeX *rx = resx_slot (h, depth);

		     *res   = X.prange;
		     res->x = (unsigned short) F.x;
		     res->y = (unsigned short) F.y;
		     rx->mv_tree_bits  = X.mvt;
		     rx->mv_coord_bits = X.mvc;
		     rx->mv_type       = (signed char) X.mctype;	/* 1 FORWARD, 2 BACKWARD, 3 INTERPOLATED */
		     rx->mv_fx	       = (signed char) (X.mctype != 2 ? X.mx : 0);
		     rx->mv_fy	       = (signed char) (X.mctype != 2 ? X.my : 0);
		     rx->mv_bx	       = (signed char) (X.mctype != 1 ? X.bx : 0);
		     rx->mv_by	       = (signed char) (X.mctype != 1 ? X.by : 0);
		     rx->prediction    = 1;
		     if (X.try_nd)		/* the predicting edge: state 0 (prediction.c:458-465) */
		     {
			res->into [0]	= 0;
			res->into [1]	= FB_NO_EDGE;
			res->weight [0] = X.ndw;
			rx->mv_type	= 0;
		     }
		     h->ret_costs = (res->tree_bits + res->matrix_bits + res->weights_bits
				     + rx->mv_tree_bits + rx->mv_coord_bits + 0.0f + 0.0f) * h->price
				    + res->err;
		     nstate = ST_RETURN;
		  }
	       }
	       else
	       {
		  /* the prediction lost: models and automaton as the first two alternatives left
		     them (prediction.c:159-186) */
		  cta_copy_s16<NT> (sh.blob, snap + 2 * P.blob_len, P.blob_len);
		  if (tid < 32)
		  {
		     unsigned *tm = (unsigned *) ((char *) h + offsetof (ShHdr, tree_counts));

		     for (int i = tid; i < 2 * FB200_MAXLEVEL; i += 32)
			tm [i] = tsnap [2 * FB200_MAXLEVEL + i];
		     __syncwarp ();
		  }
		  for (unsigned s = F.states_snap + tid; s < X.rec_states; s += NT)
		     GP (W.domain_type) [s] = GP (W.saved_dt) [s];
		  __syncthreads ();
		  if (tid == 0)
		  {
		     /* the tail of the shared pool list, which the attempt may have overwritten:
			the usable states of the split alternative, in order */
		     unsigned k = ((const unsigned short *) snap) [MB_N];

		     if ((int) k < h->pool_lo)
			h->pool_lo = (int) k;
		     for (unsigned s = F.states_snap; s < X.rec_states; s++)
			if (GP (W.domain_type) [s] & 2)
			   sh.pool [k++] = (short) s;
		     if (k != BLOB_U16 (sh, MB_N))
			h->status = FB200_ECUDA;	/* cannot happen: the lists are prefixes of each other */
		     h->states	 = X.rec_states;
		     h->hole_off = X.saved_off;
		     X.pred_done = 1;
		     nstate	 = ST_DECIDE;
		  }
	       }
	    }
	    break;
	 }

	 case ST_DECIDE:
	 {
	    const float lin = F.lincomb_costs, sub = F.subdivide_costs;
	    short      *snap  = (sh.snaps ? sh.snaps : GP (W.snap)) + (size_t) depth * SN * P.blob_len;
	    unsigned   *tsnap = sh.tsnap + (size_t) depth * TS * 2 * FB200_MAXLEVEL;

	    if (MOTION && (h->fx [depth].try_mc || h->fx [depth].try_nd) && !h->fx [depth].pred_done)
	    {
	       /* alternative 3: motion compensation + approximation of the prediction error
		  (subdivide.c:383-407, predict_range prediction.c:96-191, mc_prediction :262-370) */
	       FrameX	&X     = h->fx [depth];
	       const int level = F.level;

	       /* keep the models of the first two alternatives, hide their states, start again
		  from the models of the node's entry */
	       cta_copy_s16<NT> (snap + 2 * P.blob_len, sh.blob, P.blob_len);
	       if (tid < 32)
	       {
		  unsigned *tm = (unsigned *) ((char *) h + offsetof (ShHdr, tree_counts));

		  for (int i = tid; i < 2 * FB200_MAXLEVEL; i += 32)
		  {
		     tsnap [2 * FB200_MAXLEVEL + i] = tm [i];
		     tm [i]			    = tsnap [i];
		  }
		  __syncwarp ();
	       }
	       for (unsigned s = F.states_snap + tid; s < h->states; s += NT)
	       {
		  GP (W.saved_dt) [s]	 = GP (W.domain_type) [s];
		  GP (W.domain_type) [s] = 0;
	       }
	       __syncthreads ();
	       cta_copy_s16<NT> (sh.blob, snap, P.blob_len);
	       if (tid == 0)
	       {
		  X.rec_states = h->states;
		  X.max_pred   = fmin2 (fmin2 (lin, sub), F.max_costs);
		  /* the reference moves the states of the split alternative aside: the attempt's
		     states take their numbers (store_state_data, prediction.c:502) */
		  X.saved_off = h->hole_off;
		  h->hole_off = X.off_snap + (int) (h->states - F.states_snap);
	       }
	       __syncthreads ();
	       if (X.try_nd)
	       {
		  /* nd_prediction (prediction.c:371-400): the range's DC component, quantised with the
		     DC format, priced by the coefficient model of the node's entry and by the
		     prediction tree model */
		  if (tid == 0)
		  {
		     const float  x	 = GP (W.T) [(size_t) F.image * P.s_cap];
		     const float  y	 = GP (W.diag) [(size_t) (level - P.lmin) * P.s_cap];
		     const float  wt	 = dev_btor (dev_rtob (x / y, P.dc_m, P.dc_range), P.dc_m, P.dc_range);
		     const short *counts = sh.blob + MB_COUNTS;
		     const int	  code	 = dev_rtob (wt, P.dc_m, P.dc_range);

		     X.ndw    = wt;
		     X.mctype = 0;
		     X.mx = X.my = X.bx = X.by = 0;
		     X.mvt    = t0_ptree_bits (h, 0, level);
		     /* (a weight that rounds to zero, code RPF_ZERO = -1: the reference reads the word in
			front of its counts, coeff.c:237 -- zero, the upper end of the allocator's chunk
			size -- and gets infinitely many bits: never taken) */
		     X.mvc    = code < 0 ? FB_MAXCOSTS * FB_MAXCOSTS
					 : (float) (0.0 - dev_log2d (counts [code] / (float) sh.blob [MB_TOTALS]));
		     X.pcosts = h->price * (X.mvc + X.mvt);
		  }
	       }
	       else
	       {
		  if (level == P.p_min)
		     cta_fill_norms<NT> (P, W, F.x, F.y, level);
		  cta_find_best_mv<NT> (P, W, sh, F.x, F.y, level, h->price, 0);
		  if (tid == 0)
		  {
		     h->fi     = h->best_i;
		     h->fcosts = h->best_c;
		  }
	       }
	       if (!X.try_nd && P.motion == 2)
	       {
		  /* find_B_frame_mc (mwfa.c:341-542) without cross-B search: the best forward vector,
		     the best backward vector, and both together */
		  cta_find_best_mv<NT> (P, W, sh, F.x, F.y, level, h->price, 1);
		  if (tid == 0)
		  {
		     h->bi     = h->best_i;
		     h->bcosts = h->best_c;
		     h->isum   = -1;
		  }
		  __syncthreads ();
		  {
		     /* squared norm of the interpolated prediction error (mcpe_norm, mwfa.c:651-684):
			a sum of integers, exact in fp32 in any order while it stays below 2^24 */
		     const int	    sr = P.sr;
		     const int	    fx = h->fi < 0 ? 0 : h->fi % (2 * sr) - sr, fy = h->fi < 0 ? 0 : h->fi / (2 * sr) - sr;
		     const int	    bx = h->bi < 0 ? 0 : h->bi % (2 * sr) - sr, by = h->bi < 0 ? 0 : h->bi / (2 * sr) - sr;
		     const int	    bw = (int) width_of_level (level), bh = (int) height_of_level (level);
		     const int16_t *orig = GP (W.pix), *past = GP (W.past), *fut = GP (W.future);
		     long long	    acc = 0;

		     for (int i = tid; i < bw * bh; i += NT)
		     {
			const int x = (int) F.x + i % bw, y = (int) F.y + i / bw;
			const int ref = ((int) past [(size_t) (y + fy) * P.width + x + fx]
					 + (int) fut [(size_t) (y + by) * P.width + x + bx]) / 2;
			const int16_t d = (int16_t) (orig [(size_t) y * P.width + x] - ref);
			const int     q = d / 16;

			acc += q * q;
		     }
#pragma unroll
		     for (int o = 16; o > 0; o >>= 1)
		     {
			const unsigned lo = __shfl_xor_sync (0xffffffffu, (unsigned) (acc & 0xffffffffll), o);
			const unsigned hi = __shfl_xor_sync (0xffffffffu, (unsigned) (acc >> 32), o);

			acc += (long long) (((unsigned long long) hi << 32) | lo);
		     }
		     if ((tid & 31) == 0)
			((long long *) sh.num) [tid >> 5] = acc;
		     __syncthreads ();
		     if (tid == 0)
		     {
			long long tot = 0;
			float	  norm;

			for (int t = 0; t < NT / 32; t++)
			   tot += ((const long long *) sh.num) [t];
			if (tot <= (1ll << 24))
			   norm = (float) tot;
			else
			{
			   norm = 0;		/* the reference's order: rows, left to right */
			   for (int i = 0; i < bw * bh; i++)
			   {
			      const int x = (int) F.x + i % bw, y = (int) F.y + i / bw;
			      const int ref = ((int) past [(size_t) (y + fy) * P.width + x + fx]
					       + (int) fut [(size_t) (y + by) * P.width + x + bx]) / 2;
			      const int16_t d = (int16_t) (orig [(size_t) y * P.width + x] - ref);
			      const int	    q = d / 16;

			      norm += (float) (q * q);
			   }
			}
			const float fbits = (float) c_mv_code_length [fx + sr] + (float) c_mv_code_length [fy + sr];
			const float bbits = (float) c_mv_code_length [bx + sr] + (float) c_mv_code_length [by + sr];
			const float ibits = fbits + bbits;
			const float fc	  = h->fcosts + 3 * h->price;
			const float bc	  = h->bcosts + 3 * h->price;
			const float ic	  = norm + (ibits + 2) * h->price;

			if (fc <= ic)
			   X.mctype = fc <= bc ? 1 : 2;
			else
			   X.mctype = bc <= ic ? 2 : 3;
			X.mx = fx; X.my = fy; X.bx = bx; X.by = by;
			X.mvt = X.mctype == 3 ? 2.0f : 3.0f;
			X.mvc = X.mctype == 1 ? fbits : X.mctype == 2 ? bbits : ibits;
			X.pcosts = (X.mvt + X.mvc) * h->price;
		     }
		  }
	       }
	       else if (!X.try_nd && tid == 0)
	       {
		  /* find_P_frame_mc (mwfa.c:301-339) */
		  const int bi = h->fi < 0 ? 0 : h->fi;

		  X.mx	   = bi % (2 * P.sr) - P.sr;
		  X.my	   = bi / (2 * P.sr) - P.sr;
		  X.bx	   = X.by = 0;
		  X.mctype = 1;
		  X.mvt	   = 1.0f;
		  X.mvc	   = (float) c_mv_code_length [X.mx + P.sr] + (float) c_mv_code_length [X.my + P.sr];
		  X.pcosts = (X.mvt + X.mvc) * h->price;
	       }
	       __syncthreads ();
	       if (X.pcosts < X.max_pred)
	       {
		  /* the prediction error replaces the pixels, fresh product tables, and
		     subdivide() on it with the delta models */
		  if (X.try_nd)
		     cta_nd_range<NT> (P, W, sh, shn, F.address, level, -X.ndw * GP (W.img) [0]);
		  else
		     cta_mcpe_range<NT> (P, W, shn, F.x, F.y, level, X.mctype, X.mx, X.my, X.bx, X.by);
		  if (tid == 0)
		  {
		     float *t = W.T;
		     W.T  = W.T2;
		     W.T2 = t;
		     h->tswap ^= 1;
		     h->top	  = level;
		     h->nest_base = depth + 1;
		     X.last_state = h->states - 1;
		  }
		  __syncthreads ();
		  if (CL)
		  {
		     /* the products of the prediction-error block: every block of the cluster its share */
		     if (tid == 0)
		     {
			h->job.type   = CJ_TERR;
			h->job.states = h->states;
			h->job.level  = level;
			h->job.nested = 1;
		     }
		     cta_cluster_post<NT> (P, sh, false);
		     cta_compute_T<NT, clustered_shape<NT, MOTION> ()> (P, W, shn, 0, 0, level, level, 0,
									(unsigned) P.cluster * NT);
		  }
		  else
		     cta_compute_T<NT> (P, W, shn, 0, 0, level, level);
		  if (tid == 0)
		  {
		     Frame &C = h->frames [depth + 1];

		     C.max_costs = X.max_pred - X.pcosts;
		     C.x	 = F.x;
		     C.y	 = F.y;
		     C.image	 = 0;
		     C.address	 = 0;
		     C.gaddr	 = F.gaddr;
		     C.level	 = level;
		     C.y_state	 = F.y_state;
		     C.spec_k	 = -1;
		     h->fx [depth + 1].delta	  = 1;
		     h->fx [depth + 1].prediction = 0;
		     X.prange.into [0] = FB_NO_EDGE;
		     X.prange.tree     = FB_RANGE;
		     ndepth = depth + 1;
		     nstate = ST_ENTER;
		  }
	       }
	       else if (tid == 0)
	       {
		  /* the vector alone is too expensive: as if the nested pass had failed */
		  h->ret_costs = FB_MAXCOSTS;
		  h->nest_base = -2;	/* no tables to swap back */
		  nstate       = ST_PRED_DONE;
	       }
	       break;
	    }

	    if ((lin >= FB_MAXCOSTS && sub >= FB_MAXCOSTS) || lin < sub)
	    {
	       const bool fail = (lin >= FB_MAXCOSTS && sub >= FB_MAXCOSTS);

	       /* the models and the tree model of the node's entry -- plus, if the linear
		  combination wins, what it adds to them ("lc" models, t0_range_epilogue); drop the
		  states created below this node (subdivide.c:409-467).  Warp 0 only: thread 0 goes
		  on without a block-wide barrier in between */
	       if (tid < 32)
	       {
		  unsigned *tm = (unsigned *) ((char *) h + offsetof (ShHdr, tree_counts));

		  const short *from = EAGER && !fail ? snap + P.blob_len : snap;

		  for (int i = tid; i < P.blob_len / 8; i += 32)
		     ((uint4 *) sh.blob) [i] = ((const uint4 *) from) [i];
		  for (int i = tid; i < 2 * FB200_MAXLEVEL; i += 32)
		     tm [i] = tsnap [i];
		  __syncwarp ();
	       }
	       if (tid == 0)
	       {
		  h->states = F.states_snap;	/* remove_states (wfalib.c:276-310) */
		  if (MOTION)
		     h->hole_off = h->fx [depth].off_snap;
		  if (fail)
		     h->ret_costs = FB_MAXCOSTS;
		  else
		  {
		     if (!EAGER)
		     {
			t0_rle_update (cs, F.lrange.into, F.lc_ystate);
			t0_aac_update (P, cs, F.lc_code, F.lrange.into, F.level);
		     }
		     const unsigned short rx = (unsigned short) F.x, ry = (unsigned short) F.y;
		     *res      = F.lrange;
		     res->tree = FB_RANGE;
		     res->x    = rx;
		     res->y    = ry;
		     if (MOTION)
			*resx_slot (h, depth) = h->fx [depth].lrange;
		     h->ret_costs = lin;
		  }
		  nstate = ST_RETURN;
	       }
	    }
	    else
	    {
	       /* new state (subdivide.c:468-501, init_new_state :549-610) */
	       const int aux = band > 0
			       || F.x + width_of_level (F.level) > (unsigned) P.width
			       || F.y + height_of_level (F.level) > (unsigned) P.height;

	       if (tid == 0)
	       {
		  const unsigned s = h->states;

		  if (!aux)
		  {
		     /* rle_append (domain-pool.c:832-852); in a predicted frame the state enters
			the normal and the delta pool (subdivide.c:573-584), which share the list */
		     const unsigned n = BLOB_U16 (sh, MB_N);
		     if (n < BLOB_U16 (sh, MB_MAXDOM))
		     {
			sh.pool [n]	    = (short) s;
			if ((int) n < h->pool_lo)
			   h->pool_lo = (int) n;
			BLOB_U16 (sh, MB_N) = (unsigned short) (n + 1);
			if (MOTION)
			   BLOB_U16 (shn, MB_N) = (unsigned short) (n + 1);
		     }
		  }
		  for (int label = 0; label < 2; label++)
		  {
		     const RangeRes &c = F.child [label];

		     GP (W.tree) [2 * s + label]    = c.tree;
		     GP (W.y_state) [2 * s + label] = (short) F.new_y_state [label];
		     GP (W.x) [2 * s + label]       = c.x;
		     GP (W.y) [2 * s + label]       = c.y;
		     GP (W.y_column) [2 * s + label] = 0;
		     GP (W.into) [(size_t) (2 * s + label) * 6] = FB_NO_EDGE;
		     if (MOTION)
		     {
			const RangeX &cx = h->fx [depth].child [label];

			GP (W.mv_type) [2 * s + label] = cx.mv_type;
			GP (W.mv_fx) [2 * s + label]   = cx.mv_fx;
			GP (W.mv_fy) [2 * s + label]   = cx.mv_fy;
			GP (W.mv_bx) [2 * s + label]   = cx.mv_bx;
			GP (W.mv_by) [2 * s + label]   = cx.mv_by;
		     }
		     for (int e = 0; c.into [e] != FB_NO_EDGE; e++)
		     {
			t0_append_edge (W, s, c.into [e], c.weight [e], label);
			if (c.into [e] == F.new_y_state [label])
			   GP (W.y_column) [2 * s + label] = 1;
		     }
		     if (MOTION && W.yc_ref && band > 0)
			GP (W.yc_ref) [2 * ((int) s - h->hole_off) + label] = GP (W.y_column) [2 * s + label];
		  }
		  /* (the luminance band writes zeros only: its highest state number is enough) */
		  if (MOTION && band == 0 && s + 1 - (unsigned) h->hole_off > h->ref_max)
		     h->ref_max = s + 1 - (unsigned) h->hole_off;
		  t0_store_trans (W, s);
		  res->into [0]	    = FB_NO_EDGE;
		  res->tree	    = (short) s;
		  res->x	    = (unsigned short) F.x;
		  res->y	    = (unsigned short) F.y;
		  res->err	    = F.r_err;
		  res->tree_bits    = F.r_tree_bits;
		  res->matrix_bits  = F.r_matrix_bits;
		  res->weights_bits = F.r_weights_bits;
		  if (MOTION)
		  {
		     RangeX *rx = resx_slot (h, depth);

		     rx->mv_tree_bits  = h->fx [depth].r_mvt;
		     rx->mv_coord_bits = h->fx [depth].r_mvc;
		     rx->mv_type = rx->mv_fx = rx->mv_fy = rx->mv_bx = rx->mv_by = rx->prediction = 0;
		  }
		  h->ret_costs	    = sub;
		  nstate	    = ST_RETURN;
	       }
	       __syncthreads ();
	       const long long t0c = clock64 ();
	       cta_append_state<NT, clustered_shape<NT, MOTION> ()> (P, W, sh, aux, F.level);
	       if (tid == 0)
		  h->cyc_append += clock64 () - t0c;
	    }
	    break;
	 }
      }
      /* thread 0 continues with the scalar part of the control flow */
      if (tid == 0)
      {
	 if (clustered_shape<NT, MOTION> ())
	    LAP (h, LAP_DECIDE);	/* (diagnostics: what the case left unaccounted) */
	 t0_advance<MOTION, clustered_shape<NT, MOTION> ()> (P, W, sh, nstate, ndepth);
	 if (clustered_shape<NT, MOTION> ())
	    LAP (h, LAP_AP_STAGED);	/* (diagnostics of the clustered shape: the scalar walk) */
	 h->state [(it + 1) & 1]  = nstate;
	 h->depthv [(it + 1) & 1] = ndepth;
      }
      __syncthreads ();
   }
   __syncthreads ();
}

/*****************************************************************************
	      colour: chroma dictionary and virtual root states (thread 0)
*****************************************************************************/

/*
 *  Before the Cb band (codec/coder.c:776-800): shrink the pool to the chroma_max_states
 *  most referenced luminance states -- rle_chroma (domain-pool.c:854-879) with
 *  compute_hits (wfalib.c:182-231): state 0 first, then by hit count descending (ties in
 *  ascending state order, the order a stable sort of the reference's hit list gives),
 *  stopping at the first unreferenced state; the survivors sorted by state number -- and
 *  raise lc_min_level to the finest level the luminance used.  'hits' is block-wide
 *  scratch of at least 'states' ints.
 */
__device__ void
t0_chroma_setup (const DevParams &P, const TileWs &W, const Sh &sh, int *hits)
{
   ShHdr	 *h	 = sh.h;
   const unsigned states = h->states;
   unsigned	  max_domains = (unsigned) P.chroma_max_states;
   const unsigned n_pool = BLOB_U16 (sh, MB_N);

   if (max_domains < n_pool)
   {
      const unsigned from = 3, to = states - 1;
      unsigned	     n	  = max_domains < to ? max_domains : to;
      unsigned	     cnt  = 1;

      for (unsigned d = 0; d < to; d++)
	 hits [d] = 0;
      for (unsigned s = from; s <= to; s++)
	 for (int label = 0; label < 2; label++)
	 {
	    const short *in = GP (W.into) + (size_t) (2 * s + label) * 6;
	    for (int e = 0; in [e] != FB_NO_EDGE; e++)
	       hits [in [e]]++;
	 }
      /* selection in the order of the stable descending sort; chosen entries are
	 marked by a negative count */
      sh.pool [0] = 0;
      while (cnt < n)
      {
	 int best = -1, bestkey = 0;
	 for (unsigned d = 1; d < to; d++)
	    if (hits [d] > bestkey)
	    {
	       best    = (int) d;
	       bestkey = hits [d];
	    }
	 if (best < 0)
	    break;
	 hits [best] = -1;
	 cnt++;
      }
      /* ascending state order */
      cnt = 1;
      for (unsigned d = 1; d < to; d++)
	 if (hits [d] < 0)
	    sh.pool [cnt++] = (short) d;
      BLOB_U16 (sh, MB_N) = (unsigned short) cnt;
   }
   h->pool_lo = 0;
   BLOB_U16 (sh, MB_YINDEX) = 0;
   BLOB_U16 (sh, MB_MAXDOM) = BLOB_U16 (sh, MB_N);

   /* don't partition the chroma bands finer than the luminance band */
   {
      unsigned min_level = FB200_MAXLEVEL;

      for (unsigned s = 3; s < states; s++)
	 if (GP (W.tree) [2 * s] == FB_RANGE || GP (W.tree) [2 * s + 1] == FB_RANGE)
	 {
	    const unsigned l = (unsigned) GP (W.level_of_state) [s] - 1;
	    if (l < min_level)
	       min_level = l;
	 }
      h->lc_min = (int) min_level;
   }
   h->y_states = states;
}

/* thread 0: a virtual state whose two children are given states (coder.c:821-848) */
__device__ int
t0_virtual_state (const DevParams &P, const TileWs &W, ShHdr *h, int child0, int child1,
		  int level)
{
   const unsigned s = h->states;

   for (int label = 0; label < 2; label++)
   {
      GP (W.tree) [2 * s + label]	 = (short) (label ? child1 : child0);
      GP (W.y_state) [2 * s + label]	 = FB_RANGE;
      GP (W.x) [2 * s + label]	 = 0;
      GP (W.y) [2 * s + label]	 = 0;
      /* the reference never sets this entry for a virtual state and writes it to the stream
	 (output/matrices.c:491-516): it is what the last state that had this number left there --
	 in this frame or in the frames before it (yc_ref, TileWs) */
      GP (W.y_column) [2 * s + label] = W.yc_ref ? GP (W.yc_ref) [2 * ((int) s - h->hole_off) + label] : 0;
      GP (W.into) [(size_t) (2 * s + label) * 6] = FB_NO_EDGE;
   }
   GP (W.final_d) [s]	= t0_final_distribution (W, s);
   GP (W.level_of_state) [s] = (uint8_t) level;
   GP (W.domain_type) [s]	= 0;
   h->states		= s + 1;
   if (s + 1 >= FB200_MAXSTATES)
      h->status = FB200_EMAXSTATES;
   else if (s + 1 >= (unsigned) P.s_cap)
      h->status = FB200_ECAPACITY;
   return (int) s;
}

/*****************************************************************************
				the kernel
*****************************************************************************/

template <int NT, bool MOTION>
__global__ void __launch_bounds__ (NT, (NT >= 512 ? 1 : NT >= 256 ? 2 : NT >= 128 ? FB200_MINB : 5))
fiasco_tile_kernel (DevParams P, const TileWs *ws_array)
{
   FB_DYN_SMEM (unsigned char, smem_raw);
   /* the tile's pointer table lives in shared memory: 19 pointers are 38 registers that the
      pursuit loops need more urgently */
#ifdef FB200_EMU		/* (the emulator's "static shared" is one object for all blocks of a cluster) */
   TileWs	    &s_W  = *(TileWs *) (smem_raw + P.sm_off [17]);
#else
   __shared__ TileWs s_W;
#endif
   int		     slot = -1;
   /* cluster per stream: rank 0 encodes, the others serve it */
   constexpr bool    CLU    = clustered_shape<NT, MOTION> ();
   const unsigned    crank  = CLU && P.cluster > 1 ? cl_rank () : 0u;
   const unsigned    tile   = CLU && P.cluster > 1 ? blockIdx.x / (unsigned) P.cluster : blockIdx.x;
   const unsigned    ntiles = CLU && P.cluster > 1 ? gridDim.x / (unsigned) P.cluster : gridDim.x;

   if (threadIdx.x == 0)
   {
      s_W = ws_array [tile];
      /* pipelined batches (fb200_encode_tiles): the tile's pixels are on their way; the copy that
	 follows them in the same stream writes the flag */
      if (P.tile_ready)
      {
#ifdef FB200_EMU
	 (void) 0;		/* (the emulated copies are synchronous) */
#else
	 int v;

	 do
	    asm volatile ("ld.global.acquire.gpu.b32 %0, [%1];" : "=r" (v) : "l" (P.tile_ready + tile) : "memory");
	 while (v != P.ready_epoch);
#endif
      }
   }

   /*
    *  More tiles than workspaces: take a free one (entry i of ws_array also describes
    *  workspace i).  At most n_slots blocks are resident at any time, so a free one exists;
    *  the scan starts at a different place on every SM.  Everything the kernel reads from a
    *  workspace it has written itself before, so the previous user's data never shows.
    */
   if ((int) ntiles > P.n_slots)	/* uniform; never with a cluster (launcher) */
   {
      __shared__ int s_slot;

      if (threadIdx.x == 0)
      {
	 unsigned smid = 0;
#ifndef FB200_EMU
	 asm volatile ("mov.u32 %0, %%smid;" : "=r" (smid));
#endif
	 int i = (int) ((smid * 4u) % (unsigned) P.n_slots);
	 while (atomicCAS (P.slot_flags + i, 0, 1) != 0)
	    i = i + 1 == P.n_slots ? 0 : i + 1;
	 __threadfence ();
	 s_slot = i;
      }
      __syncthreads ();
      slot = s_slot;
      if (threadIdx.x == 0)
      {
	 const TileWs &S = ws_array [slot];
	 s_W.img   = S.img;
	 s_W.T	   = S.T;
	 s_W.SS	   = S.SS;
	 s_W.diag  = S.diag;
	 s_W.trans = S.trans;
	 s_W.Gglob = S.Gglob;
	 s_W.snap  = S.snap;
	 if (MOTION)
	 {
	    s_W.T2	 = S.T2;
	    s_W.norms	 = S.norms;
	    s_W.pix2	 = S.pix2;
	    s_W.norm2	 = S.norm2;
	    s_W.saved_dt = S.saved_dt;
	 }
      }
   }
   __syncthreads ();
   const TileWs &W  = s_W;
   const Sh	sh  = carve (smem_raw, P, NT, GP (W.Gglob) + (size_t) crank * FB_MAXEDGES * (P.s_cap + 1));
   const Sh    &sh_base = sh;
   ShHdr       *h   = sh.h;
   const int	tid = threadIdx.x;

   if (tid == 0)
   {
      h->status	   = FB200_OK;
      h->frames	   = sh.frames;
      h->fx	   = sh.fx;
      h->nest_base = -1;
      h->w.l2_dc   = sh.l2;
      h->w.l2_lv   = sh.l2 + P.aac_dc_size;
      h->trace_len = 0;
      h->mp_calls = h->mp_steps = h->pass2 = h->blocks = h->ip_bytes = 0;
      h->mp_bytes = h->ss_bytes = 0;
      h->cyc_T = h->cyc_mp = h->cyc_append = 0;
      h->cyc_start = clock64 ();
      h->lap_last  = h->cyc_start;
      for (int i = 0; i < 16; i++)
	 h->lap [i] = 0;
      h->states	   = 0;
      h->pool_lo   = 0;
      h->spec_len  = 0;
      h->tswap	   = 0;
      h->mbar_phase = 0;
      mbar_init (&h->mbar);
   }
   __syncthreads ();

   /* quantiser value tables and the codes of the pass-1 dummy weight 0.5 */
   for (int i = tid; i < P.aac_dc_size + P.aac_lvl_size; i += NT)
      sh.qt_dc [i] = i < P.aac_dc_size ? dev_btor (i, P.dc_m, P.dc_range)
				       : dev_btor (i - P.aac_dc_size, P.rpf_m, P.rpf_range);
   if (tid == 0)
   {
      h->w.half_lv = (short) dev_rtob (0.5f, P.rpf_m, P.rpf_range);
      h->w.half_dc = (short) dev_rtob (0.5f, P.dc_m, P.dc_range);
   }
   __syncthreads ();

   if (CLU && crank != 0)
   {
      cta_helper_loop<NT, MOTION> (P, s_W, sh, crank);
      return;
   }

   {
      cta_init_basis<NT> (P, W, sh);
      /* init_tree_model (bintree.c:70-93), rle_model_alloc (domain-pool.c:655-672),
	 aac_model_alloc (coeff.c:285-313) */
      if (tid == 0)
      {
	 const unsigned c0 [FB200_MAXLEVEL] = {20, 17, 15, 10, 5, 4, 3, 2, 1, 1, 1,
					       1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
	 const unsigned c1 [FB200_MAXLEVEL] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 3, 5,
					       10, 15, 20, 25, 30, 35, 60, 60, 60, 60};
	 for (int l = 0; l < FB200_MAXLEVEL; l++)
	 {
	    h->tree_counts [l] = MOTION ? c1 [l] * 0x10001u : c1 [l];
	    h->tree_total [l]  = MOTION ? (c0 [l] + c1 [l]) * 0x10001u : c0 [l] + c1 [l];
	 }
      }
      for (int i = tid; i < P.blob_len; i += NT)
	 sh.blob [i] = (i % P.blob_half) >= MB_COUNTS - 1 ? 1 : 0;
      if (MOTION)
	 for (int i = tid; i < 2 * P.s_cap; i += NT)
	    GP (W.mv_type) [i] = GP (W.mv_fx) [i] = GP (W.mv_fy) [i] = GP (W.mv_bx) [i] = GP (W.mv_by) [i] = 0;
      __syncthreads ();
      /* a predicted frame has a second model set for the prediction errors: the delta pool and
	 the delta coefficient model (coder.c:713-736), initialised like the first */
      for (int set = 0; set < (MOTION ? 2 : 1); set++)
      if (tid == 0)
      {
	 Sh sh = sh_base;

	 sh.blob += set * P.blob_half;
	 for (int k = 0; k < FB_MAXEDGES + 1; k++)
	    BLOB_S16 (sh, MB_COUNT + k) = 1;
	 BLOB_U16 (sh, MB_TOTAL)  = FB_MAXEDGES + 1;
	 BLOB_U16 (sh, MB_MAXDOM) = (unsigned short) P.max_domains;
	 /* the basis states that may be used as domains: 0, 1, 2 */
	 for (unsigned s = 0; s < 3; s++)
	    if (BLOB_U16 (sh, MB_N) < BLOB_U16 (sh, MB_MAXDOM))
	    {
	       sh.pool [BLOB_U16 (sh, MB_N)] = (short) s;
	       BLOB_U16 (sh, MB_N)++;
	       if (s == 0)
	       {
		  BLOB_S16 (sh, MB_D0INDEX) = 0;
		  BLOB_U16 (sh, MB_D0N)	    = 1;
	       }
	    }
	 sh.blob [MB_TOTALS] = (short) P.aac_dc_size;
	 for (int l = P.lc_min; l <= P.lc_max; l++)
	    sh.blob [MB_TOTALS + l - P.coeff_min_level + 1] = (short) P.aac_lvl_size;
      }
      __syncthreads ();
   }

   /* the bands of the frame, one subdivide() pass each (coder.c:738-849) */
   int band_tree [3] = {FB_RANGE, FB_RANGE, FB_RANGE};
   int ycb_node	     = -1;

   if (tid == 0)
   {
      /* a frame of a colour sequence starts with the range levels the chroma bands of the frame
	 before left behind (coder.c:797: c->options.lc_min_level is never set back) */
      const int given = P.tile_lc_min ? P.tile_lc_min [tile] : 0;

      h->lc_min = given > P.lc_min && given <= P.lc_max ? given : P.lc_min;
   }
   __syncthreads ();
   if (tid == 0)
   {
      h->hole_off = 0;
      h->ref_max  = 0;
   }
   for (int band = 0; band < P.bands && h->status == FB200_OK; band++)
   {
      if (band == 1)
      {
	 if (tid == 0)
	    t0_chroma_setup (P, W, sh, (int *) sh.bnd);
	 __syncthreads ();
	 /* every state number the luminance band has used holds a zero now (control.c:190) */
	 if (MOTION && W.yc_ref)
	    for (unsigned i = tid; i < 2 * h->ref_max; i += NT)
	       GP (W.yc_ref) [i] = 0;
	 /* a predicted colour frame: the motion compensation of the luminance tree comes off the
	    chroma planes before they are coded (coder.c:798-799) */
	 if (MOTION && (P.motion == 1 || P.motion == 2))
	    cta_subtract_mc<NT> (P, W, h->states);
      }
      cta_subdivide_band<NT, MOTION> (P, s_W, sh, band, band ? band_tree [0] : FB_RANGE);
      if (h->status != FB200_OK)
	 break;
      band_tree [band] = h->root.tree;
      if (tid == 0)
      {
	 TileResult *r = W.result;

	 if (h->root.tree == FB_RANGE)
	    h->status = FB200_ENOROOT;
	 r->costs [band]	= h->ret_costs;
	 r->err [band]		= h->root.err;
	 r->tree_bits [band]	= h->root.tree_bits;
	 r->matrix_bits [band]	= h->root.matrix_bits;
	 r->weights_bits [band] = h->root.weights_bits;
	 for (int i = 0; i < 4; i++)
	    r->progress [band][i] = h->progress [i];
	 if (band == 1 && h->status == FB200_OK)
	    h->w.index = t0_virtual_state (P, W, h, band_tree [0], band_tree [1], P.level + 1);
      }
      __syncthreads ();
      if (band == 1)
	 ycb_node = h->w.index;
   }
   if (P.bands == 3 && tid == 0 && h->status == FB200_OK)
   {
      const int cr = t0_virtual_state (P, W, h, band_tree [2], FB_RANGE, P.level + 1);
      if (h->status == FB200_OK)
	 h->root.tree = (short) t0_virtual_state (P, W, h, ycb_node, cr, P.level + 2);
   }
   __syncthreads ();

   if (tid == 0)
   {
      TileResult *r = W.result;

      r->status	      = h->status;
      r->states	      = h->states;
      r->basis_states = 3;
      r->root_state   = h->root.tree >= 0 ? (unsigned) h->root.tree : 0;
      r->trace_len = h->trace_len;
      r->lc_min_end = h->lc_min;
      r->mp_calls  = h->mp_calls;
      r->mp_steps  = h->mp_steps;
      r->pass2	   = h->pass2;
      r->blocks	   = h->blocks;
      r->ip_bytes  = h->ip_bytes;
      r->mp_bytes  = h->mp_bytes;
      r->ss_bytes  = h->ss_bytes;
      r->cyc_total  = (unsigned long long) (clock64 () - h->cyc_start);
      r->cyc_T	    = (unsigned long long) h->cyc_T;
      r->cyc_mp	    = (unsigned long long) h->cyc_mp;
      r->cyc_append = (unsigned long long) h->cyc_append;
      for (int i = 0; i < 16; i++)
	 r->lap [i] = (unsigned long long) h->lap [i];
   }
   if (slot >= 0)
   {
      __syncthreads ();
      if (tid == 0)
      {
	 __threadfence ();
	 atomicExch (P.slot_flags + slot, 0);
      }
   }
   if (CLU && P.cluster > 1)
   {
      if (tid == 0)
      {
	 h->job.type   = CJ_EXIT;
	 h->job.nested = 0;
      }
      cta_cluster_post<NT> (P, sh, false);
   }
}

/*****************************************************************************
			       probe kernel
*****************************************************************************/

__device__ __forceinline__ float range_of_enum (int e)
{
   return e == 0 ? 0.75f : e == 2 ? 1.5f : e == 3 ? 2.0f : 1.0f;
}

__global__ void
fiasco_probe_kernel (int kind, int n, const float *f, const int *a, const int *b,
		     const int *c, int *out_i, float *out_f)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;

   if (i >= n)
      return;
   switch (kind)
   {
      case 0:
	 out_i [i] = dev_rtob (f [i], a [i], range_of_enum (b [i]));
	 break;
      case 1:
	 out_f [i] = dev_btor (a [i], b [i], range_of_enum (c [i]));
	 break;
      case 2:
	 out_i [i] = (int) dev_bits_bin_code ((unsigned) a [i], (unsigned) b [i]);
	 break;
      case 3:
	 out_f [i] = neg_log2f_via_double (a [i] / (float) b [i]);
	 break;
   }
}

} /* namespace */

/*****************************************************************************
				host launchers
*****************************************************************************/

int
fb_tile_kernel_threads (const DevParams &p, int n_tiles)
{
   /* threads span the domain pool: small tiles have ~100-250 domains, a 1024^2 frame
      up to ~1500 */
   const char *e = getenv ("FB200_NT");	/* experiments only */
   if (e && (atoi (e) == 96 || atoi (e) == 128 || atoi (e) == 256 || atoi (e) == 512))
      return atoi (e);
   (void) p;
   /* few tiles: latency matters, give each tile a whole SM's worth of threads; a full
      batch: 128 threads per tile and several tiles resident per SM hide each other's
      serial phases ("thread-block shape") */
   int sms = 148, dev = 0;
   if (cudaGetDevice (&dev) == cudaSuccess)
      cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev);
   return n_tiles < sms ? 512 : 128;
}

size_t
fb_tile_kernel_smem (const DevParams &p, int nt)
{
   size_t off [20];

   return smem_layout (p, nt, off);
}

static bool g_tables_ready [64];

static cudaError_t
upload_tables (void)
{
   int dev = 0;
   cudaError_t e = cudaGetDevice (&dev);
   if (e != cudaSuccess)
      return e;
   if (dev >= 0 && dev < 64 && g_tables_ready [dev])
      return cudaSuccess;
   /* init_matrix_probabilities (domain-pool.c:970-999) */
   static float m0 [1024], m1 [1024];
   unsigned	index = 0;
   for (unsigned n = 1; n <= 9; n++)
      for (unsigned ex = 0; ex < 1u << n; ex++, index++)
      {
	 m1 [index] = (float) -log2 ((double) (1 / (float) (1 << n)));
	 m0 [index] = (float) -log2 ((double) (1 - 1 / (float) (1 << n)));
      }
   e = cudaMemcpyToSymbol (c_matrix_0, m0, sizeof m0);
   if (e != cudaSuccess)
      return e;
   e = cudaMemcpyToSymbol (c_matrix_1, m1, sizeof m1);
   if (e != cudaSuccess)
      return e;
   if (dev >= 0 && dev < 64)
      g_tables_ready [dev] = true;
   return cudaSuccess;
}

/* the 512-thread shape has an SM to itself: Gram-Schmidt rows and model snapshots on chip
   whenever they fit (the batch shape keeps them in global memory to fit four blocks per SM) */
static void
shape_params (DevParams &p, int nt)
{
   if (nt >= 512 && p.big && !getenv ("FB200_BIG"))
   {
      DevParams q = p;
      size_t	off [20];

      q.big = 0;
      if (smem_layout (q, nt, off) <= 200 * 1024)
	 p.big = 0;
   }
}

template <int NT, bool MOTION>
static cudaError_t
launch_nt (const DevParams &p_in, const TileWs *d_ws, int n_tiles, cudaStream_t stream, int cluster = 1)
{
   const auto kernel = fiasco_tile_kernel<NT, MOTION>;
   DevParams	p = p_in;
   size_t off [20];

   shape_params (p, NT);
   const size_t smem = smem_layout (p, NT, off);

   for (int i = 0; i < 20; i++)
      p.sm_off [i] = off [i] == (size_t) -1 ? 0xffffffffu : (unsigned) off [i];
   cudaError_t	e    = cudaFuncSetAttribute (kernel,
					     cudaFuncAttributeMaxDynamicSharedMemorySize,
					     (int) smem);
   if (e != cudaSuccess)
      return e;
   {
      const char *cv = getenv ("FB200_CARVE");	/* experiments only: shared-memory carve-out in % */
      if (cv)
	 cudaFuncSetAttribute (kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi (cv));
   }
   p.cluster = cluster > 1 ? cluster : 1;
   if (p.cluster > 1)
   {
#ifdef FB200_EMU
      FB_LAUNCH_CLUSTER (kernel, n_tiles * p.cluster, NT, p.cluster, smem, stream, p, d_ws);
#else
      cudaLaunchConfig_t  cfg = {};
      cudaLaunchAttribute at [1];

      cfg.gridDim	   = dim3 ((unsigned) (n_tiles * p.cluster));
      cfg.blockDim	   = dim3 (NT);
      cfg.dynamicSmemBytes = smem;
      cfg.stream	   = stream;
      at [0].id		       = cudaLaunchAttributeClusterDimension;
      at [0].val.clusterDim.x = (unsigned) p.cluster;
      at [0].val.clusterDim.y = 1;
      at [0].val.clusterDim.z = 1;
      cfg.attrs	   = at;
      cfg.numAttrs = 1;
      return cudaLaunchKernelEx (&cfg, kernel, p, d_ws);
#endif
   }
   else
      FB_LAUNCH (kernel, n_tiles, NT, smem, stream, p, d_ws);
   return cudaGetLastError ();
}

cudaError_t
fb_launch_tile_kernel (const DevParams &p, const TileWs *d_ws, int n_tiles,
		       cudaStream_t stream)
{
   cudaError_t e = upload_tables ();
   if (e != cudaSuccess)
      return e;
   if (p.motion)		/* predicted frames: the two production shapes only */
      return fb_tile_kernel_threads (p, n_tiles) <= 128
	     ? launch_nt<128, true> (p, d_ws, n_tiles, stream)
	     : launch_nt<512, true> (p, d_ws, n_tiles, stream, fb_tile_kernel_cluster (p, n_tiles));
   switch (fb_tile_kernel_threads (p, n_tiles))
   {
      case 96:	return launch_nt<96, false> (p, d_ws, n_tiles, stream);
      case 128: return launch_nt<128, false> (p, d_ws, n_tiles, stream);
      case 256: return launch_nt<256, false> (p, d_ws, n_tiles, stream);
      default:	return launch_nt<512, false> (p, d_ws, n_tiles, stream, fb_tile_kernel_cluster (p, n_tiles));
   }
}

/*
 *  Thread blocks per stream.  A full batch keeps every SM busy with streams of its own; with
 *  fewer streams than SMs the idle SMs join the streams as helper blocks (powers of two up to
 *  the portable cluster size 8): a single 1024^2 frame runs on 8 SMs, the 64 tiles of a
 *  4096^2 frame on 128.
 */
int
fb_tile_kernel_cluster (const DevParams &p, int n_tiles)
{
   int sms = 148, dev = 0, c = 1;

   if (n_tiles <= 0 || n_tiles > p.n_slots)
      return 1;
   if (fb_tile_kernel_threads (p, n_tiles) != 512)
      return 1;
   if (cudaGetDevice (&dev) == cudaSuccess)
      cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev);
   while (c * 2 <= FB_MAXCLUSTER && n_tiles * c * 2 <= sms)
      c *= 2;
   {
      const char *e = getenv ("FB200_CLUSTER");	/* experiments / tests: force the size */
      if (e && atoi (e) >= 1 && atoi (e) <= FB_MAXCLUSTER)
	 c = atoi (e);
   }
#ifndef FB200_EMU
   /* as many clusters as streams must be able to run at once (they are co-scheduled per GPC) */
   while (c > 1)
   {
      cudaLaunchConfig_t  cfg = {};
      cudaLaunchAttribute at [1];
      int		  n = 0;
      DevParams		  q = p;
      size_t		  off [20];

      shape_params (q, 512);
      cfg.gridDim	   = dim3 ((unsigned) (n_tiles * c));
      cfg.blockDim	   = dim3 (512);
      cfg.dynamicSmemBytes = smem_layout (q, 512, off);
      at [0].id		       = cudaLaunchAttributeClusterDimension;
      at [0].val.clusterDim.x = (unsigned) c;
      at [0].val.clusterDim.y = at [0].val.clusterDim.z = 1;
      cfg.attrs	   = at;
      cfg.numAttrs = 1;
      const auto kernel = p.motion ? fiasco_tile_kernel<512, true> : fiasco_tile_kernel<512, false>;

      cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) cfg.dynamicSmemBytes);
      if (cudaOccupancyMaxActiveClusters (&n, kernel, &cfg) == cudaSuccess && n >= n_tiles)
	 break;
      cudaGetLastError ();
      c /= 2;
   }
#endif
   return c;
}

template <int NT>
static int
occupancy_nt (const DevParams &p)
{
   int	  n    = 0;
   size_t smem = fb_tile_kernel_smem (p, NT);

   const auto kernel = p.motion && (NT == 128 || NT == 512)
			? fiasco_tile_kernel<(NT == 128 ? 128 : 512), true>
			: fiasco_tile_kernel<NT, false>;
   if (cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
			     (int) smem) != cudaSuccess
       || cudaOccupancyMaxActiveBlocksPerMultiprocessor (&n, kernel, NT, smem)
	  != cudaSuccess)
   {
      cudaGetLastError ();
      return 0;
   }
   return n;
}

int
fb_tile_kernel_occupancy (const DevParams &p)
{
   switch (fb_tile_kernel_threads (p, 1 << 20))
   {
      case 96:	return occupancy_nt<96> (p);
      case 128: return occupancy_nt<128> (p);
      case 256: return occupancy_nt<256> (p);
      default:	return occupancy_nt<512> (p);
   }
}

cudaError_t
fb_launch_probe (int kind, int n, const float *f, const int *a, const int *b,
		 const int *c, int *out_i, float *out_f)
{
   FB_LAUNCH (fiasco_probe_kernel, (n + 255) / 256, 256, 0, 0, kind, n, f, a, b, c, out_i, out_f);
   return cudaGetLastError ();
}
