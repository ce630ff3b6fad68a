/*
 *  ffi.cu -- the extern "C" boundary of include/fiasco_b200.h: parameter clamping,
 *  device workspace, H2D / launch / D2H, error mapping.  No CPU fallback exists: without
 *  a CUDA device every compute entry point fails with FB200_ENODEVICE.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#include "tile_kernel.cuh"

struct fb200_ctx
{
   fb200_params_t params;
   fb200_motion_t motion;	/* frame_type 0: a context for intra frames */
   int16_t	 *d_past;	/* [tiles][w*h] reference frames (predicted frames only) */
   int16_t	 *d_future;	/* [tiles][w*h] backward references (B frames only) */
   DevParams	  dp;
   int		  max_tiles, device, nt;
   size_t	  smem;
   /* device memory */
   unsigned char *d_work;	/* [n_slots] private tables of the tiles in flight */
   size_t	  work_stride;
   int		  n_slots;	/* workspaces: min (max_tiles, tiles that can be resident at once) */
   int		 *d_slot_flags;	/* [n_slots] 0 = free; taken and released by the kernel's blocks */
   int16_t	 *d_pix;	/* [tiles][bands][w*h] */
   size_t	  pix_elems;	/* per tile */
   unsigned char *d_wfa;	/* [tiles][wfa_block] */
   size_t	  wfa_block;
   TileResult	 *d_results;
   fb200_trace_rec_t *d_trace;	/* tile 0 only */
   int		  trace_cap;
   TileWs	 *d_ws;
   int		 *d_lc_min;	/* [tiles] lc_min_level the tiles start with (fb200_wfa_t.lc_min_level) */
   uint8_t	 *d_yc;		/* [tiles][2 * FB200_MAXSTATES] colour sequences: fb200_wfa_t.y_column_history */
   int		 *d_ready;	/* [tiles] epoch of the launch whose pixels have arrived (pipelined batches) */
   int		 *h_epoch;	/* pinned: the current epoch, source of the flag copies */
   int		  epoch;
   cudaStream_t	  copy_stream;
   /* pinned host staging */
   unsigned char *h_wfa;
   TileResult	 *h_results;
   std::vector<TileWs> ws;
   cudaStream_t	  stream;
   cudaEvent_t	  ev [6];
   fb200_stats_t  stats;
   int		  launched_tiles;
   cudaStream_t	  last_stream;
   int		  auto_grow;	/* capacity was chosen by us: grow and retry on overflow */
};

static void
set_err (char *err, size_t errlen, const char *fmt, ...)
{
   va_list ap;

   if (!err || !errlen)
      return;
   va_start (ap, fmt);
   vsnprintf (err, errlen, fmt, ap);
   va_end (ap);
}

#define CUDA_TRY(call)                                                              \
   do {                                                                             \
      cudaError_t e_ = (call);                                                      \
      if (e_ != cudaSuccess) {                                                      \
	 set_err (err, errlen, "CUDA error %s at %s:%d (%s)", cudaGetErrorName (e_), \
		  __FILE__, __LINE__, cudaGetErrorString (e_));                     \
	 return e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver        \
		? FB200_ENODEVICE : FB200_ECUDA;                                    \
      }                                                                             \
   } while (0)

/* process-wide work counters (fb200_counters_get): what the launches of this process have done
   since the last reset, whichever entry point -- fiasco_coder() included -- made them */
#include <mutex>
static std::mutex	g_counters_lock;
static fb200_counters_t g_counters;

static void
count_work (double kernel_ms, uint64_t launches, uint64_t h2d, uint64_t d2h)
{
   std::lock_guard<std::mutex> hold (g_counters_lock);

   g_counters.kernel_ms += kernel_ms;
   g_counters.launches	+= launches;
   g_counters.h2d_bytes += h2d;
   g_counters.d2h_bytes += d2h;
}

extern "C" void
fb200_counters_get (fb200_counters_t *out)
{
   std::lock_guard<std::mutex> hold (g_counters_lock);

   if (out)
      *out = g_counters;
}

extern "C" void
fb200_counters_reset (void)
{
   std::lock_guard<std::mutex> hold (g_counters_lock);

   memset (&g_counters, 0, sizeof g_counters);
}

extern "C" const char *
fb200_version (void)
{
   return "fiasco_b200 0.1 (sm_100a persistent tile kernel)";
}

extern "C" int
fb200_device_count (void)
{
   int n = 0;

   if (cudaGetDeviceCount (&n) != cudaSuccess)
   {
      cudaGetLastError ();
      return 0;
   }
   return n;
}

static float
range_value (int e)		/* lib/rpf.c:202-222 */
{
   switch (e)
   {
      case 0:  return 0.75f;
      case 2:  return 1.5f;
      case 3:  return 2.0f;
      default: return 1.0f;
   }
}

static unsigned
image_level (unsigned width, unsigned height)	/* codec/coder.c:249-256 */
{
   unsigned lx = (unsigned) (log2 ((double) (width - 1)) + 1);
   unsigned ly = (unsigned) (log2 ((double) (height - 1)) + 1);

   return (lx > ly ? lx : ly) * 2 - ((ly == lx + 1) ? 1 : 0);
}

extern "C" int
fb200_params_init (fb200_params_t *p, int width, int height, int bands, float quality,
		   int optimize, char *err, size_t errlen)
{
   int lc_min, lc_max, max_elements;

   if (!p)
      return FB200_EINVAL;
   memset (p, 0, sizeof *p);
   if (width < 2 || height < 2 || (width & 1) || (height & 1))
   {
      /* lib/image.c:194-195 */
      set_err (err, errlen, "Width and height of images must be even numbers.");
      return FB200_EINVAL;
   }
   if (quality <= 0)
   {
      set_err (err, errlen, "Compression quality has to be positive."); /* coder.c:118 */
      return FB200_EINVAL;
   }
   if (bands != 1 && bands != 3)
   {
      set_err (err, errlen, "bands must be 1 (grey) or 3 (YCbCr)");
      return FB200_EINVAL;
   }
   if (optimize >= 3)
   {
      set_err (err, errlen, "optimization level >= 3 (full search) is not supported: "
	       "the reference has undefined behaviour there");
      return FB200_EUNSUPPORTED;
   }
   /* CLI mapping, bin/cwfa.c:326-345 */
   if (optimize <= 0)
   {
      lc_max = 10; lc_min = 6; max_elements = 3; optimize = 0;
   }
   else
   {
      lc_max = 12; lc_min = 4; max_elements = 5; optimize -= 1;
   }
   p->width  = width;
   p->height = height;
   p->bands  = bands;
   p->level  = (int) image_level ((unsigned) width, (unsigned) height);
   /* alloc_coder, codec/coder.c:260-296 (tiling exponent is always 0) */
   p->lc_min_level = lc_min > 3 ? lc_min : 3;
   p->lc_max_level = lc_max < p->level - 1 ? lc_max : p->level - 1;
   if (p->lc_min_level > p->lc_max_level)
      p->lc_min_level = p->lc_max_level;
   p->images_level = 5 < p->lc_max_level - 1 ? 5 : p->lc_max_level - 1;
   p->max_elements = max_elements;
   p->max_states   = FB200_MAXSTATES;	/* min (10000, MAXSTATES), coder.c:326 */
   p->chroma_max_states = 40;
   p->price	      = 128 * 64 / quality;
   p->chroma_decrease = 2.0f;
   p->rpf_mantissa    = 3;
   p->rpf_range	      = range_value (2);
   p->dc_rpf_mantissa = 5;
   p->dc_rpf_range    = range_value (1);
   p->second_domain_block = optimize > 0;
   p->state_capacity  = 0;
   return FB200_OK;
}

static int
default_capacity (const fb200_params_t *p)
{
   /* measured on the synthetic frames at q=20: 236 states for 256^2, 481 for 512^2,
      1487 for 1024^2, 4755 for 2048^2 -- roughly 1.1-3.7 per 32x32 block */
   const double blocks = ((double) p->width * p->height) / 1024.0;
   int		cap    = (int) (128 + 1.75 * blocks);

   cap = (cap + 63) / 64 * 64;
   if (cap > FB200_MAXSTATES)
      cap = FB200_MAXSTATES;
   return cap;
}

static int
derive (const fb200_params_t *p, const fb200_motion_t *mo, DevParams *d, char *err, size_t errlen)
{
   memset (d, 0, sizeof *d);
   if (p->level < 1 || p->level > FB200_MAXLEVEL || p->lc_max_level > 12
       || p->lc_max_level < p->lc_min_level || p->lc_min_level < 3
       || p->images_level < 1 || p->images_level > 5
       || p->images_level >= p->lc_max_level + 1
       || p->max_elements < 1 || p->max_elements > FB_MAXEDGES
       || p->rpf_mantissa < 2 || p->rpf_mantissa > 8
       || p->dc_rpf_mantissa < 2 || p->dc_rpf_mantissa > 8)
   {
      set_err (err, errlen, "parameters outside the supported range");
      return FB200_EINVAL;
   }
   if (p->bands != 1 && p->bands != 3)
   {
      set_err (err, errlen, "bands must be 1 (grey) or 3 (Y, Cb, Cr)");
      return FB200_EINVAL;
   }
   d->width  = p->width;
   d->height = p->height;
   d->level  = p->level;
   d->bands  = p->bands;
   d->lc_min = p->lc_min_level;
   d->lc_max = p->lc_max_level;
   d->il     = p->images_level;
   d->lmin   = d->lc_min < d->il ? d->lc_min : d->il;
   d->nlev   = d->lc_max - d->lmin + 1;
   d->tn     = (1 << d->nlev) - 1;
   d->max_elements = p->max_elements;
   d->max_domains  = p->max_states < FB200_MAXSTATES ? p->max_states : FB200_MAXSTATES;
   if (d->max_domains < 1)
      d->max_domains = 1;
   d->chroma_max_states = p->chroma_max_states > 1 ? p->chroma_max_states : 1;
   d->price	      = p->price;
   d->chroma_decrease = p->chroma_decrease;
   d->rpf_m	= p->rpf_mantissa;
   d->dc_m	= p->dc_rpf_mantissa;
   d->rpf_range = p->rpf_range;
   d->dc_range	= p->dc_rpf_range;
   d->second_domain_block = p->second_domain_block;
   d->s_cap = p->state_capacity > 0 ? p->state_capacity : default_capacity (p);
   if (mo && mo->frame_type)
   {
      if (mo->frame_type < 1 || mo->frame_type > FB200_FRAME_INTRA || (p->bands != 1 && p->bands != 3)
	  || (p->bands == 3 && mo->frame_type == FB200_FRAME_ND) || mo->search_range < 1
	  || mo->search_range > 16)
      {
	 set_err (err, errlen, "predicted frames: P and B frames (grey or colour), or grey intra frames "
		  "with nondeterministic prediction; search range 1..16");
	 return FB200_EUNSUPPORTED;
      }
      /* prediction levels are a subset of the range levels (coder.c:284-290) */
      d->motion = mo->frame_type;
      d->p_min	= mo->p_min_level > d->lc_min ? mo->p_min_level : d->lc_min;
      d->p_max	= mo->p_max_level < d->lc_max ? mo->p_max_level : d->lc_max;
      if (d->p_min > d->p_max)
	 d->p_min = d->p_max;
      d->sr = mo->search_range;
      /* states of losing alternatives stay behind as holes until the host closes them: ~3x the
	 final states while a frame is built (2773 on BASELINE config 5, whose
	 frames end with ~930); the motion search also borrows 2 x 512 floats of the pursuit's
	 work arrays.  Too small a guess costs a second launch (FB200_ECAPACITY, ctx_grow). */
      if (p->state_capacity <= 0)
	 d->s_cap = 4 * d->s_cap;
      if (d->s_cap < 512)
	 d->s_cap = 512;
   }
   if (d->s_cap < 8)
      d->s_cap = 8;
   if (d->s_cap > FB200_MAXSTATES)
      d->s_cap = FB200_MAXSTATES;
   d->s_cap = (d->s_cap + 7) / 8 * 8;
   d->coeff_min_level = d->lc_min;
   d->aac_dc_size  = 1 << (1 + d->dc_m);
   d->aac_lvl_size = 1 << (1 + d->rpf_m);
   d->blob_len	   = MB_COUNTS + d->aac_dc_size
		     + (d->lc_max - d->lc_min + 1) * d->aac_lvl_size;
   d->blob_len	   = (d->blob_len + 7) / 8 * 8;
   d->blob_half	   = d->blob_len;
   d->n_frames	   = d->level - d->lc_min + 2;
   if (d->motion)
   {
      d->blob_len  = 2 * d->blob_half;	/* normal and delta model sets */
      d->n_frames += d->p_max - d->lc_min + 2;	/* nested pass over the prediction error */
   }
   d->big	   = d->s_cap > 768 ? 3 : 0;
   {
      const char *e = getenv ("FB200_BIG");	/* experiments only */
      if (e)
	 d->big = atoi (e);
   }
   return FB200_OK;
}

static size_t
up256 (size_t x)
{
   return (x + 255) & ~(size_t) 255;
}

/* carve the private tables of one tile out of d_work */
static size_t
work_layout (const DevParams &d, size_t *off /* [12] */)
{
   size_t o = 0, sc = (size_t) d.s_cap;

   off [0] = o; o += up256 (sc * FB_IMG_STRIDE * 4);			/* img */
   off [1] = o; o += up256 ((size_t) d.tn * sc * 4);			/* T */
   off [2] = o; o += up256 ((size_t) d.nlev * sc * sc * 4);		/* SS */
   off [3] = o; o += up256 ((size_t) d.nlev * sc * 4);			/* diag */
   off [4] = o; o += up256 ((size_t) (d.n_frames > FB_MAXDEPTH ? d.n_frames : FB_MAXDEPTH)
			   * 3 * d.blob_len * 2);				/* snap */
   off [5] = o; o += up256 (sc * sizeof (Trans));			/* trans */
   off [6] = o; o += up256 ((sc + 1) * 4 * FB_MAXEDGES) * FB_MAXCLUSTER;	/* Gglob, per block of a cluster */
   for (int i = 7; i < 12; i++)
      off [i] = o;
   if (d.motion)
   {
      off [7]  = o; o += up256 ((size_t) d.tn * sc * 4);			/* T2 */
      off [8]  = o; o += up256 ((size_t) (d.motion == 2 ? 2 : 1) * (d.p_max - d.p_min + 1)
				* 4 * d.sr * d.sr * 4);				/* norms */
      off [9]  = o; o += up256 (((size_t) 1 << d.lc_max) * 4);		/* pix2 */
      off [10] = o; o += up256 ((size_t) d.tn * 4);				/* norm2 */
      off [11] = o; o += up256 (sc);						/* saved_dt */
   }
   return o;
}

/* the automaton of one tile, contiguous so that it travels in one copy */
static size_t
wfa_layout (const DevParams &d, size_t *off /* [15] */)
{
   size_t o = 0, sc = (size_t) d.s_cap;

   off [0] = o; o += up256 (sc * 4);		/* final_distribution */
   off [1] = o; o += up256 (sc * 2 * 6 * 4);	/* weight */
   off [2] = o; o += up256 (sc * 2 * 6 * 2);	/* into */
   off [3] = o; o += up256 (sc * 2 * 2);	/* tree */
   off [4] = o; o += up256 (sc * 2 * 2);	/* x */
   off [5] = o; o += up256 (sc * 2 * 2);	/* y */
   off [6] = o; o += up256 (sc * 2 * 2);	/* y_state */
   off [7] = o; o += up256 (sc);		/* level_of_state */
   off [8] = o; o += up256 (sc);		/* domain_type */
   off [9] = o; o += up256 (sc * 2);		/* y_column */
   off [10] = off [11] = off [12] = off [13] = off [14] = o;
   if (d.motion)
   {
      off [10] = o; o += up256 (sc * 2);	/* mv_type */
      off [11] = o; o += up256 (sc * 2);	/* mv_fx */
      off [12] = o; o += up256 (sc * 2);	/* mv_fy */
      off [13] = o; o += up256 (sc * 2);	/* mv_bx */
      off [14] = o; o += up256 (sc * 2);	/* mv_by */
   }
   return o;
}

extern "C" void
fb200_destroy (fb200_ctx_t *c)
{
   if (!c)
      return;
   cudaSetDevice (c->device);
   cudaFree (c->d_work);
   cudaFree (c->d_slot_flags);
   cudaFree (c->d_pix);
   cudaFree (c->d_past);
   cudaFree (c->d_future);
   cudaFree (c->d_wfa);
   cudaFree (c->d_results);
   cudaFree (c->d_trace);
   cudaFree (c->d_ws);
   cudaFree (c->d_lc_min);
   cudaFree (c->d_yc);
   cudaFree (c->d_ready);
   cudaFreeHost (c->h_epoch);
   if (c->copy_stream)
      cudaStreamDestroy (c->copy_stream);
   cudaFreeHost (c->h_wfa);
   cudaFreeHost (c->h_results);
   for (int i = 0; i < 6; i++)
      if (c->ev [i])
	 cudaEventDestroy (c->ev [i]);
   if (c->stream)
      cudaStreamDestroy (c->stream);
   delete c;
}

static int
ctx_alloc (fb200_ctx_t *c, char *err, size_t errlen)
{
   const DevParams &d	   = c->dp;
   const int	    device = c->device, max_tiles = c->max_tiles;

   CUDA_TRY (cudaSetDevice (device));
   c->nt   = 512;
   c->smem = fb_tile_kernel_smem (d, c->nt);
   {
      int max_smem = 0;
      cudaDeviceGetAttribute (&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
      if ((size_t) max_smem < c->smem)
      {
	 set_err (err, errlen, "state capacity %d needs %zu bytes of shared memory, device "
		  "offers %d", d.s_cap, c->smem, max_smem);
	 return FB200_EINVAL;
      }
   }
   size_t woff [12], aoff [15];
   c->work_stride = work_layout (d, woff);
   c->wfa_block	  = wfa_layout (d, aoff);
   c->pix_elems	  = (size_t) d.bands * d.width * d.height;

   /* The big tables (88 MB per 1024^2 frame) exist once per tile IN FLIGHT, not once per tile
      of the batch: a block takes a free workspace when it starts and hands it back when its
      frame is done, so a launch may hold any number of tiles (several waves even out the
      different running times of the frames). */
   {
      int sms = 0;
      cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, device);
      const int resident = sms * fb_tile_kernel_occupancy (d);
      c->n_slots = (resident > 0 && resident < max_tiles) ? resident : max_tiles;
   }
   CUDA_TRY (cudaMalloc (&c->d_work, c->work_stride * c->n_slots));
   cudaFree (c->d_slot_flags);
   c->d_slot_flags = NULL;
   CUDA_TRY (cudaMalloc (&c->d_slot_flags, sizeof (int) * c->n_slots));
   CUDA_TRY (cudaMemset (c->d_slot_flags, 0, sizeof (int) * c->n_slots));
   c->dp.n_slots    = c->n_slots;
   c->dp.slot_flags = c->d_slot_flags;
   CUDA_TRY (cudaMalloc (&c->d_wfa, c->wfa_block * max_tiles));
   c->h_wfa = NULL;		/* pinned staging of the automata: allocated by the first download, so that
				   a context can be set up beside host threads that fault pages in */
   if (!c->d_pix)
   {
      CUDA_TRY (cudaMalloc (&c->d_pix, c->pix_elems * 2 * max_tiles));
      if (d.motion)
	 CUDA_TRY (cudaMalloc (&c->d_past, c->pix_elems * 2 * max_tiles));
      if (d.motion == 2)
	 CUDA_TRY (cudaMalloc (&c->d_future, c->pix_elems * 2 * max_tiles));
      CUDA_TRY (cudaMalloc (&c->d_results, sizeof (TileResult) * max_tiles));
      CUDA_TRY (cudaMalloc (&c->d_ws, sizeof (TileWs) * max_tiles));
      if (d.motion && d.bands == 3)
      {
	 CUDA_TRY (cudaMalloc (&c->d_yc, (size_t) 2 * FB200_MAXSTATES * max_tiles));
	 CUDA_TRY (cudaMemset (c->d_yc, 0, (size_t) 2 * FB200_MAXSTATES * max_tiles));
      }
      CUDA_TRY (cudaMalloc (&c->d_lc_min, sizeof (int) * max_tiles));
      CUDA_TRY (cudaMemset (c->d_lc_min, 0, sizeof (int) * max_tiles));
      CUDA_TRY (cudaMalloc (&c->d_ready, sizeof (int) * max_tiles));
      CUDA_TRY (cudaMemset (c->d_ready, 0, sizeof (int) * max_tiles));
      CUDA_TRY (cudaMallocHost (&c->h_epoch, sizeof (int)));
      CUDA_TRY (cudaStreamCreateWithFlags (&c->copy_stream, cudaStreamNonBlocking));
      CUDA_TRY (cudaMallocHost (&c->h_results, sizeof (TileResult) * max_tiles));
      CUDA_TRY (cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking));
      for (int i = 0; i < 6; i++)
	 CUDA_TRY (cudaEventCreate (&c->ev [i]));
   }
   c->ws.resize (max_tiles);
   for (int t = 0; t < max_tiles; t++)
   {
      /* tile t's entry also describes workspace t (for t < n_slots) */
      unsigned char *wb = c->d_work + c->work_stride * (t < c->n_slots ? t : 0);
      unsigned char *ab = c->d_wfa + c->wfa_block * t;
      TileWs	    &w	= c->ws [t];

      memset (&w, 0, sizeof w);
      w.pix	  = c->d_pix + c->pix_elems * t;
      w.img	  = (float *) (wb + woff [0]);
      w.T	  = (float *) (wb + woff [1]);
      w.SS	  = (float *) (wb + woff [2]);
      w.diag	  = (float *) (wb + woff [3]);
      w.snap	  = (int16_t *) (wb + woff [4]);
      w.trans	  = (Trans *) (wb + woff [5]);
      w.Gglob	  = (float *) (wb + woff [6]);
      w.final_d	       = (float *) (ab + aoff [0]);
      w.weight	       = (float *) (ab + aoff [1]);
      w.into	       = (int16_t *) (ab + aoff [2]);
      w.tree	       = (int16_t *) (ab + aoff [3]);
      w.x	       = (uint16_t *) (ab + aoff [4]);
      w.y	       = (uint16_t *) (ab + aoff [5]);
      w.y_state	       = (int16_t *) (ab + aoff [6]);
      w.level_of_state = (uint8_t *) (ab + aoff [7]);
      w.domain_type    = (uint8_t *) (ab + aoff [8]);
      w.y_column       = (uint8_t *) (ab + aoff [9]);
      if (d.motion)
      {
	 w.past	    = c->d_past + c->pix_elems * t;
	 w.future   = d.motion == 2 ? c->d_future + c->pix_elems * t : w.past;
	 w.mv_bx    = (int8_t *) (ab + aoff [13]);
	 w.mv_by    = (int8_t *) (ab + aoff [14]);
	 w.T2	    = (float *) (wb + woff [7]);
	 w.norms    = (float *) (wb + woff [8]);
	 w.pix2	    = (float *) (wb + woff [9]);
	 w.norm2    = (int *) (wb + woff [10]);
	 w.saved_dt = (uint8_t *) (wb + woff [11]);
	 w.mv_type  = (int8_t *) (ab + aoff [10]);
	 w.mv_fx    = (int8_t *) (ab + aoff [11]);
	 w.mv_fy    = (int8_t *) (ab + aoff [12]);
	 w.yc_ref   = c->d_yc ? c->d_yc + (size_t) 2 * FB200_MAXSTATES * t : NULL;
      }
      w.result	       = c->d_results + t;
      w.trace	       = t == 0 ? c->d_trace : NULL;
   }
   CUDA_TRY (cudaMemcpy (c->d_ws, c->ws.data (), sizeof (TileWs) * max_tiles,
			 cudaMemcpyHostToDevice));
   return FB200_OK;
}

/* grow the per-tile state capacity (after FB200_ECAPACITY); pixels stay resident */
static int
ctx_grow (fb200_ctx_t *c, char *err, size_t errlen)
{
   fb200_params_t p = c->params;
   DevParams	  d;
   int		  cap = c->dp.s_cap + c->dp.s_cap / 2;

   if (c->dp.s_cap >= FB200_MAXSTATES)
      return FB200_ECAPACITY;
   if (cap > FB200_MAXSTATES)
      cap = FB200_MAXSTATES;
   p.state_capacity = cap;
   int rc = derive (&p, &c->motion, &d, err, errlen);
   if (rc)
      return rc;
   d.trace_cap	 = c->dp.trace_cap;
   d.tile_lc_min = c->dp.tile_lc_min;
   CUDA_TRY (cudaSetDevice (c->device));
   cudaFree (c->d_work);
   cudaFree (c->d_wfa);
   cudaFreeHost (c->h_wfa);
   c->d_work = NULL;
   c->d_wfa  = NULL;
   c->h_wfa  = NULL;
   c->dp     = d;
   c->params.state_capacity = cap;
   return ctx_alloc (c, err, errlen);
}

static int
create_ctx (fb200_ctx_t **out, const fb200_params_t *p, const fb200_motion_t *mo, int max_tiles,
	    int device, char *err, size_t errlen);

extern "C" int
fb200_create (fb200_ctx_t **out, const fb200_params_t *p, int max_tiles, int device,
	      char *err, size_t errlen)
{
   return create_ctx (out, p, NULL, max_tiles, device, err, errlen);
}

extern "C" int
fb200_create_predicted (fb200_ctx_t **out, const fb200_params_t *p, const fb200_motion_t *motion,
			int max_tiles, int device, char *err, size_t errlen)
{
   if (!motion || motion->frame_type == 0)
   {
      set_err (err, errlen, "fb200_create_predicted: a frame type is required");
      return FB200_EINVAL;
   }
   return create_ctx (out, p, motion, max_tiles, device, err, errlen);
}

static int
create_ctx (fb200_ctx_t **out, const fb200_params_t *p, const fb200_motion_t *mo, int max_tiles,
	    int device, char *err, size_t errlen)
{
   if (!out || !p || max_tiles < 1)
   {
      set_err (err, errlen, "fb200_create: bad arguments");
      return FB200_EINVAL;
   }
   *out = NULL;
   DevParams d;
   int	     rc = derive (p, mo, &d, err, errlen);
   if (rc)
      return rc;
   if (fb200_device_count () <= device)
   {
      set_err (err, errlen, "no CUDA device %d available (this library has no CPU path)",
	       device);
      return FB200_ENODEVICE;
   }
   fb200_ctx_t *c = new fb200_ctx_t ();
   c->params	  = *p;
   if (mo)
      c->motion	  = *mo;
   c->dp	  = d;
   c->max_tiles	  = max_tiles;
   c->device	  = device;
   c->trace_cap	  = 0;
   c->auto_grow	  = p->state_capacity <= 0;
   memset (&c->stats, 0, sizeof c->stats);
   rc = ctx_alloc (c, err, errlen);
   if (rc)
   {
      fb200_destroy (c);
      return rc;
   }
   *out = c;
   return FB200_OK;
}

static int
ensure_trace (fb200_ctx_t *c, int trace_cap, char *err, size_t errlen)
{
   if (trace_cap <= c->trace_cap)
      return FB200_OK;
   CUDA_TRY (cudaSetDevice (c->device));
   cudaFree (c->d_trace);
   c->d_trace = NULL;
   CUDA_TRY (cudaMalloc (&c->d_trace, sizeof (fb200_trace_rec_t) * (size_t) trace_cap));
   c->trace_cap	   = trace_cap;
   c->ws [0].trace = c->d_trace;
   CUDA_TRY (cudaMemcpy (c->d_ws, c->ws.data (), sizeof (TileWs), cudaMemcpyHostToDevice));
   return FB200_OK;
}

extern "C" int
fb200_upload (fb200_ctx_t *c, int n_tiles, const int16_t *const *planes, char *err,
	      size_t errlen)
{
   if (!c || n_tiles < 1 || n_tiles > c->max_tiles || !planes)
   {
      set_err (err, errlen, "fb200_upload: bad arguments");
      return FB200_EINVAL;
   }
   CUDA_TRY (cudaSetDevice (c->device));
   const size_t plane = (size_t) c->dp.width * c->dp.height;
   /* straight from the caller's planes: a true asynchronous DMA when they are pinned, the
      driver's own staging when they are pageable -- no second copy through a buffer of ours */
   CUDA_TRY (cudaEventRecord (c->ev [0], c->stream));
   for (int t = 0; t < n_tiles; t++)
      for (int b = 0; b < c->dp.bands; b++)
	 CUDA_TRY (cudaMemcpyAsync (c->d_pix + c->pix_elems * t + plane * b,
				    planes [t * c->dp.bands + b], plane * 2,
				    cudaMemcpyHostToDevice, c->stream));
   CUDA_TRY (cudaEventRecord (c->ev [1], c->stream));
   CUDA_TRY (cudaStreamSynchronize (c->stream));
   cudaEventElapsedTime (&c->stats.h2d_ms, c->ev [0], c->ev [1]);
   c->stats.h2d_bytes = c->pix_elems * 2 * n_tiles;
   count_work (0, 0, c->stats.h2d_bytes, 0);
   return FB200_OK;
}

extern "C" int
fb200_launch (fb200_ctx_t *c, int n_tiles, void *cuda_stream, char *err, size_t errlen)
{
   if (!c || n_tiles < 1 || n_tiles > c->max_tiles)
   {
      set_err (err, errlen, "fb200_launch: bad arguments");
      return FB200_EINVAL;
   }
   CUDA_TRY (cudaSetDevice (c->device));
   cudaStream_t s = cuda_stream ? (cudaStream_t) cuda_stream : c->stream;
   CUDA_TRY (cudaEventRecord (c->ev [2], s));
   if (n_tiles > c->n_slots)	/* a launch that died may have left workspaces marked busy */
      CUDA_TRY (cudaMemsetAsync (c->d_slot_flags, 0, sizeof (int) * c->n_slots, s));
   CUDA_TRY (fb_launch_tile_kernel (c->dp, c->d_ws, n_tiles, s));
   CUDA_TRY (cudaEventRecord (c->ev [3], s));
   c->launched_tiles	    = n_tiles;
   c->last_stream	    = s;
   c->stats.kernel_launches = 1;
   return FB200_OK;
}

extern "C" int
fb200_sync (fb200_ctx_t *c, char *err, size_t errlen)
{
   if (!c)
      return FB200_EINVAL;
   CUDA_TRY (cudaSetDevice (c->device));
   CUDA_TRY (cudaStreamSynchronize (c->last_stream ? c->last_stream : c->stream));
   if (c->launched_tiles)
      cudaEventElapsedTime (&c->stats.kernel_ms, c->ev [2], c->ev [3]);
   return FB200_OK;
}

static const char *
status_text (int st)
{
   switch (st)
   {
      case FB200_ECAPACITY:  return "state capacity of the device workspace exceeded";
      case FB200_EMAXSTATES: return "Maximum number of states reached!";
      case FB200_ENOROOT:    return "No root state generated!";
      default:		     return "tile failed";
   }
}

extern "C" int
fb200_download (fb200_ctx_t *c, int n_tiles, fb200_wfa_t *out, fb200_trace_rec_t *trace,
		int trace_cap, int *trace_len, char *err, size_t errlen)
{
   if (!c || n_tiles < 1 || n_tiles > c->max_tiles || !out)
   {
      set_err (err, errlen, "fb200_download: bad arguments");
      return FB200_EINVAL;
   }
   CUDA_TRY (cudaSetDevice (c->device));
   cudaStream_t s = c->last_stream ? c->last_stream : c->stream;
   CUDA_TRY (cudaEventRecord (c->ev [4], s));
   CUDA_TRY (cudaMemcpyAsync (c->h_results, c->d_results, sizeof (TileResult) * n_tiles,
			      cudaMemcpyDeviceToHost, s));
   if (!c->h_wfa)
      CUDA_TRY (cudaMallocHost (&c->h_wfa, c->wfa_block * c->max_tiles));
   CUDA_TRY (cudaMemcpyAsync (c->h_wfa, c->d_wfa, c->wfa_block * n_tiles,
			      cudaMemcpyDeviceToHost, s));
   CUDA_TRY (cudaEventRecord (c->ev [5], s));
   CUDA_TRY (cudaStreamSynchronize (s));
   cudaEventElapsedTime (&c->stats.kernel_ms, c->ev [2], c->ev [3]);
   cudaEventElapsedTime (&c->stats.d2h_ms, c->ev [4], c->ev [5]);
   c->stats.d2h_bytes = (sizeof (TileResult) + c->wfa_block) * n_tiles;
   count_work (c->stats.kernel_ms, 1, 0, c->stats.d2h_bytes);

   size_t aoff [15];
   wfa_layout (c->dp, aoff);
   int rc = FB200_OK;
   c->stats.ip_bytes = c->stats.mp_calls = c->stats.mp_steps = c->stats.pass2 = 0;
   c->stats.blocks = c->stats.states = 0;
   c->stats.mp_bytes = c->stats.ss_bytes = 0;
   c->stats.cyc_total = c->stats.cyc_T = c->stats.cyc_mp = c->stats.cyc_append = 0;
   memset (c->stats.lap, 0, sizeof c->stats.lap);
   for (int t = 0; t < n_tiles; t++)
   {
      const TileResult	  &r  = c->h_results [t];
      const unsigned char *ab = c->h_wfa + c->wfa_block * t;
      fb200_wfa_t	  &o  = out [t];

      o.status	     = r.status;
      o.states	     = r.states;
      o.basis_states = r.basis_states;
      o.root_state   = r.root_state;
      o.lc_min_level = r.lc_min_end;
      memcpy (o.progress, r.progress, sizeof o.progress);
      for (int b = 0; b < 3; b++)
      {
	 o.costs [b]	    = r.costs [b];
	 o.err [b]	    = r.err [b];
	 o.tree_bits [b]    = r.tree_bits [b];
	 o.matrix_bits [b]  = r.matrix_bits [b];
	 o.weights_bits [b] = r.weights_bits [b];
      }
      c->stats.ip_bytes += r.ip_bytes;
      c->stats.mp_calls += r.mp_calls;
      c->stats.mp_steps += r.mp_steps;
      c->stats.pass2	+= r.pass2;
      c->stats.blocks	+= r.blocks;
      c->stats.states	+= r.states;
      c->stats.mp_bytes += r.mp_bytes;
      c->stats.ss_bytes += r.ss_bytes;
      c->stats.cyc_total  += r.cyc_total;
      c->stats.cyc_T	  += r.cyc_T;
      c->stats.cyc_mp	  += r.cyc_mp;
      c->stats.cyc_append += r.cyc_append;
      for (int i = 0; i < 16; i++)
	 c->stats.lap [i] += r.lap [i];
      if (r.status != FB200_OK)
      {
	 if (rc == FB200_OK)
	 {
	    rc = r.status;
	    set_err (err, errlen, "tile %d: %s", t, status_text (r.status));
	 }
	 continue;
      }
      if ((unsigned) o.capacity < r.states)
      {
	 if (rc == FB200_OK)
	 {
	    rc = FB200_EINVAL;
	    set_err (err, errlen, "tile %d: output arrays hold %d states, %u needed", t,
		     o.capacity, r.states);
	 }
	 continue;
      }
      const size_t n = r.states;
      if (o.final_distribution) memcpy (o.final_distribution, ab + aoff [0], n * 4);
      if (o.weight)		memcpy (o.weight, ab + aoff [1], n * 2 * 6 * 4);
      if (o.into)		memcpy (o.into, ab + aoff [2], n * 2 * 6 * 2);
      if (o.tree)		memcpy (o.tree, ab + aoff [3], n * 2 * 2);
      if (o.x)			memcpy (o.x, ab + aoff [4], n * 2 * 2);
      if (o.y)			memcpy (o.y, ab + aoff [5], n * 2 * 2);
      if (o.y_state)		memcpy (o.y_state, ab + aoff [6], n * 2 * 2);
      if (o.level_of_state)	memcpy (o.level_of_state, ab + aoff [7], n);
      if (o.domain_type)	memcpy (o.domain_type, ab + aoff [8], n);
      if (o.y_column)		memcpy (o.y_column, ab + aoff [9], n * 2);
      if (c->dp.motion)
      {
	 if (o.mv_type)		memcpy (o.mv_type, ab + aoff [10], n * 2);
	 if (o.mv_fx)		memcpy (o.mv_fx, ab + aoff [11], n * 2);
	 if (o.mv_fy)		memcpy (o.mv_fy, ab + aoff [12], n * 2);
	 if (o.mv_bx)		memcpy (o.mv_bx, ab + aoff [13], n * 2);
	 if (o.mv_by)		memcpy (o.mv_by, ab + aoff [14], n * 2);
      }
   }

   if (trace_len)
      *trace_len = 0;
   if (trace && trace_cap > 0 && c->d_trace)
   {
      int n = c->h_results [0].trace_len;

      if (trace_len)
	 *trace_len = n;
      if (n > trace_cap)
	 n = trace_cap;
      if (n > c->trace_cap)
	 n = c->trace_cap;
      if (n > 0)
	 CUDA_TRY (cudaMemcpy (trace, c->d_trace, sizeof (fb200_trace_rec_t) * (size_t) n,
			       cudaMemcpyDeviceToHost));
   }
   return rc;
}

extern "C" int
fb200_encode_tiles (fb200_ctx_t *c, int n_tiles, const int16_t *const *planes,
		    fb200_wfa_t *out, fb200_trace_rec_t *trace, int trace_cap,
		    int *trace_len, char *err, size_t errlen)
{
   int rc;

   if (!c)
      return FB200_EINVAL;
   if (trace && trace_cap > 0)
   {
      if ((rc = ensure_trace (c, trace_cap, err, errlen)))
	 return rc;
      c->dp.trace_cap = c->trace_cap;
   }
   else
      c->dp.trace_cap = 0;
   if (n_tiles < 1 || n_tiles > c->max_tiles || !out)
   {
      set_err (err, errlen, "fb200_encode_tiles: bad arguments");
      return FB200_EINVAL;
   }
   {
      /* the range levels the tiles start with, when a caller chains frames (colour sequences) */
      std::vector<int> lc (n_tiles);
      bool		any = false;

      for (int t = 0; t < n_tiles; t++)
	 any |= (lc [t] = out [t].lc_min_level) > 0;
      c->dp.tile_lc_min = any ? c->d_lc_min : NULL;
      if (any)
      {
	 CUDA_TRY (cudaSetDevice (c->device));
	 CUDA_TRY (cudaMemcpy (c->d_lc_min, lc.data (), sizeof (int) * n_tiles, cudaMemcpyHostToDevice));
      }
   }
   /*
    *  More tiles than the device holds at once: launch first, copy after.  The blocks of the launch
    *  wait for their tile's flag, which travels behind the tile's pixels in the copy stream, so the
    *  later tiles' pixels cross the bus while the first ones are coded (FIASCO_PIPELINE=0: the plain
    *  order upload, launch).  The flag is the launch's epoch: nothing to clear between launches.
    */
   bool pipelined = n_tiles > c->n_slots && !(getenv ("FIASCO_PIPELINE") && atoi (getenv ("FIASCO_PIPELINE")) == 0);
#ifdef FB200_EMU
   pipelined = false;
#endif
   if (pipelined)
   {
      const size_t plane = (size_t) c->dp.width * c->dp.height;

      CUDA_TRY (cudaSetDevice (c->device));
      *c->h_epoch	= ++c->epoch;
      c->dp.tile_ready	= c->d_ready;
      c->dp.ready_epoch = c->epoch;
      rc = fb200_launch (c, n_tiles, NULL, err, errlen);
      c->dp.tile_ready = NULL;
      if (rc)
	 return rc;
      CUDA_TRY (cudaEventRecord (c->ev [0], c->copy_stream));
      for (int t = 0; t < n_tiles; t++)
      {
	 for (int b = 0; b < c->dp.bands; b++)
	    CUDA_TRY (cudaMemcpyAsync (c->d_pix + c->pix_elems * t + plane * b, planes [t * c->dp.bands + b],
				       plane * 2, cudaMemcpyHostToDevice, c->copy_stream));
	 CUDA_TRY (cudaMemcpyAsync (c->d_ready + t, c->h_epoch, sizeof (int), cudaMemcpyHostToDevice,
				    c->copy_stream));
      }
      CUDA_TRY (cudaEventRecord (c->ev [1], c->copy_stream));
      CUDA_TRY (cudaStreamSynchronize (c->copy_stream));
      cudaEventElapsedTime (&c->stats.h2d_ms, c->ev [0], c->ev [1]);
      c->stats.h2d_bytes = c->pix_elems * 2 * n_tiles;
      count_work (0, 0, c->stats.h2d_bytes, 0);
   }
   else if ((rc = fb200_upload (c, n_tiles, planes, err, errlen)))
      return rc;
   float total_ms = 0;
   int	 launches = 0;
   for (;;)
   {
      if (!pipelined && (rc = fb200_launch (c, n_tiles, NULL, err, errlen)))
	 return rc;
      pipelined = false;		/* (a capacity retry finds the pixels in place) */
      rc = fb200_download (c, n_tiles, out, trace, trace_cap, trace_len, err, errlen);
      /* a launch that ran out of state capacity is part of the bill */
      total_ms += c->stats.kernel_ms;
      c->stats.kernel_ms       = total_ms;
      c->stats.kernel_launches = ++launches;
      if (rc != FB200_ECAPACITY || !c->auto_grow)
	 return rc;
      /* a tile outgrew the workspace we sized ourselves: enlarge it and run again */
      if ((rc = ctx_grow (c, err, errlen)))
	 return rc;
   }
}

extern "C" int
fb200_encode_predicted (fb200_ctx_t *c, int n_tiles, const int16_t *const *planes,
			const int16_t *const *past, const int16_t *const *future,
			fb200_wfa_t *out, char *err, size_t errlen)
{
   int rc;

   if (!c || !c->dp.motion || (!past && c->dp.motion != FB200_FRAME_ND && c->dp.motion != FB200_FRAME_INTRA)
       || n_tiles < 1
       || n_tiles > c->max_tiles || (c->dp.motion == 2 && !future))
   {
      set_err (err, errlen, "fb200_encode_predicted: needs a context of fb200_create_predicted() "
	       "and one reference frame per tile (two for B frames)");
      return FB200_EINVAL;
   }
   c->dp.trace_cap = 0;
   {
      /* the range levels the frames start with (colour sequences: fb200_wfa_t.lc_min_level) */
      std::vector<int> lc (n_tiles);
      bool		any = false;

      for (int t = 0; t < n_tiles; t++)
	 any |= (lc [t] = out [t].lc_min_level) > 0;
      c->dp.tile_lc_min = any ? c->d_lc_min : NULL;
      if (any)
      {
	 CUDA_TRY (cudaSetDevice (c->device));
	 CUDA_TRY (cudaMemcpy (c->d_lc_min, lc.data (), sizeof (int) * n_tiles, cudaMemcpyHostToDevice));
      }
   }
   if ((rc = fb200_upload (c, n_tiles, planes, err, errlen)))
      return rc;
   for (int t = 0; t < n_tiles && c->dp.motion != FB200_FRAME_ND && c->dp.motion != FB200_FRAME_INTRA; t++)
      {
      CUDA_TRY (cudaMemcpyAsync (c->d_past + c->pix_elems * t, past [t], c->pix_elems * 2,
				 cudaMemcpyHostToDevice, c->stream));
      if (c->dp.motion == 2)
	 CUDA_TRY (cudaMemcpyAsync (c->d_future + c->pix_elems * t, future [t], c->pix_elems * 2,
				    cudaMemcpyHostToDevice, c->stream));
   }
   CUDA_TRY (cudaStreamSynchronize (c->stream));
   c->stats.h2d_bytes += c->pix_elems * 2 * n_tiles * (c->dp.motion == 2 ? 2 : c->dp.motion == 1 ? 1 : 0);
   count_work (0, 0, c->pix_elems * 2 * n_tiles * (c->dp.motion == 2 ? 2 : c->dp.motion == 1 ? 1 : 0), 0);
   float total_ms = 0;
   int	 launches = 0;
   for (;;)
   {
      /* colour frames: what the frames before left in the reference's y_column array (again before
	 a second launch: the first one has written to it) */
      for (int t = 0; t < n_tiles && c->d_yc; t++)
	 if (out [t].y_column_history)
	    CUDA_TRY (cudaMemcpyAsync (c->d_yc + (size_t) 2 * FB200_MAXSTATES * t, out [t].y_column_history,
				       (size_t) 2 * FB200_MAXSTATES, cudaMemcpyHostToDevice, c->stream));
	 else
	    CUDA_TRY (cudaMemsetAsync (c->d_yc + (size_t) 2 * FB200_MAXSTATES * t, 0,
				       (size_t) 2 * FB200_MAXSTATES, c->stream));
      if ((rc = fb200_launch (c, n_tiles, NULL, err, errlen)))
	 return rc;
      rc = fb200_download (c, n_tiles, out, NULL, 0, NULL, err, errlen);
      if (rc == FB200_OK && c->d_yc)
	 for (int t = 0; t < n_tiles; t++)
	    if (out [t].y_column_history)
	       CUDA_TRY (cudaMemcpy (out [t].y_column_history, c->d_yc + (size_t) 2 * FB200_MAXSTATES * t,
				     (size_t) 2 * FB200_MAXSTATES, cudaMemcpyDeviceToHost));
      total_ms += c->stats.kernel_ms;
      c->stats.kernel_ms       = total_ms;
      c->stats.kernel_launches = ++launches;
      if (rc != FB200_ECAPACITY || !c->auto_grow)
	 return rc;
      if ((rc = ctx_grow (c, err, errlen)))
	 return rc;
   }
}

extern "C" int
fb200_resident_tiles (const fb200_ctx_t *c)
{
   int sms = 0;

   if (!c || cudaSetDevice (c->device) != cudaSuccess)
      return 0;
   cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, c->device);
   return sms * fb_tile_kernel_occupancy (c->dp);
}

extern "C" int
fb200_state_capacity (const fb200_ctx_t *c)
{
   return c ? c->dp.s_cap : 0;
}

extern "C" void
fb200_get_stats (const fb200_ctx_t *c, fb200_stats_t *stats)
{
   if (c && stats)
      *stats = c->stats;
}

extern "C" int
fb200_wfa_alloc (fb200_wfa_t *w, int capacity)
{
   if (!w || capacity < 1)
      return FB200_EINVAL;
   memset (w, 0, sizeof *w);
   w->capacity		 = capacity;
   w->final_distribution = (float *) calloc (capacity, 4);
   w->level_of_state	 = (uint8_t *) calloc (capacity, 1);
   w->domain_type	 = (uint8_t *) calloc (capacity, 1);
   w->tree		 = (int16_t *) calloc ((size_t) capacity * 2, 2);
   w->x			 = (uint16_t *) calloc ((size_t) capacity * 2, 2);
   w->y			 = (uint16_t *) calloc ((size_t) capacity * 2, 2);
   w->into		 = (int16_t *) calloc ((size_t) capacity * 12, 2);
   w->weight		 = (float *) calloc ((size_t) capacity * 12, 4);
   w->y_state		 = (int16_t *) calloc ((size_t) capacity * 2, 2);
   w->y_column		 = (uint8_t *) calloc ((size_t) capacity * 2, 1);
   w->mv_type		 = (int8_t *) calloc ((size_t) capacity * 2, 1);
   w->mv_fx		 = (int8_t *) calloc ((size_t) capacity * 2, 1);
   w->mv_fy		 = (int8_t *) calloc ((size_t) capacity * 2, 1);
   w->mv_bx		 = (int8_t *) calloc ((size_t) capacity * 2, 1);
   w->mv_by		 = (int8_t *) calloc ((size_t) capacity * 2, 1);
   w->y_column_history	 = (uint8_t *) calloc ((size_t) 2 * FB200_MAXSTATES, 1);
   if (!w->y_column_history || !w->final_distribution || !w->level_of_state || !w->domain_type || !w->tree
       || !w->x || !w->y || !w->into || !w->weight || !w->y_state || !w->y_column
       || !w->mv_type || !w->mv_fx || !w->mv_fy || !w->mv_bx || !w->mv_by)
   {
      fb200_wfa_free (w);
      return FB200_EINVAL;
   }
   return FB200_OK;
}

extern "C" void
fb200_wfa_free (fb200_wfa_t *w)
{
   if (!w)
      return;
   free (w->final_distribution);
   free (w->level_of_state);
   free (w->domain_type);
   free (w->tree);
   free (w->x);
   free (w->y);
   free (w->into);
   free (w->weight);
   free (w->y_state);
   free (w->y_column);
   free (w->mv_type);
   free (w->mv_fx);
   free (w->mv_fy);
   free (w->mv_bx);
   free (w->mv_by);
   free (w->y_column_history);
   memset (w, 0, sizeof *w);
}

extern "C" int
fb200_probe (int kind, int n, const float *f, const int *a, const int *b, const int *c,
	     int *out_i, float *out_f, char *err, size_t errlen)
{
   if (n < 1 || kind < 0 || kind > 3)
      return FB200_EINVAL;
   if (fb200_device_count () < 1)
   {
      set_err (err, errlen, "no CUDA device available (this library has no CPU path)");
      return FB200_ENODEVICE;
   }
   float *df = NULL, *dof = NULL;
   int	 *da = NULL, *db = NULL, *dc = NULL, *doi = NULL;
   const size_t nb = (size_t) n * 4;

   CUDA_TRY (cudaMalloc (&df, nb));
   CUDA_TRY (cudaMalloc (&da, nb));
   CUDA_TRY (cudaMalloc (&db, nb));
   CUDA_TRY (cudaMalloc (&dc, nb));
   CUDA_TRY (cudaMalloc (&doi, nb));
   CUDA_TRY (cudaMalloc (&dof, nb));
   CUDA_TRY (cudaMemset (df, 0, nb));
   CUDA_TRY (cudaMemset (da, 0, nb));
   CUDA_TRY (cudaMemset (db, 0, nb));
   CUDA_TRY (cudaMemset (dc, 0, nb));
   if (f) CUDA_TRY (cudaMemcpy (df, f, nb, cudaMemcpyHostToDevice));
   if (a) CUDA_TRY (cudaMemcpy (da, a, nb, cudaMemcpyHostToDevice));
   if (b) CUDA_TRY (cudaMemcpy (db, b, nb, cudaMemcpyHostToDevice));
   if (c) CUDA_TRY (cudaMemcpy (dc, c, nb, cudaMemcpyHostToDevice));
   CUDA_TRY (fb_launch_probe (kind, n, df, da, db, dc, doi, dof));
   CUDA_TRY (cudaDeviceSynchronize ());
   if (out_i) CUDA_TRY (cudaMemcpy (out_i, doi, nb, cudaMemcpyDeviceToHost));
   if (out_f) CUDA_TRY (cudaMemcpy (out_f, dof, nb, cudaMemcpyDeviceToHost));
   cudaFree (df); cudaFree (da); cudaFree (db); cudaFree (dc); cudaFree (doi); cudaFree (dof);
   return FB200_OK;
}
