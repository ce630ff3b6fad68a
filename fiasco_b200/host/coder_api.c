/*
 *  coder_api.c -- fiasco_coder(): the public entry point of the encoder.
 *
 *  What the reference does in codec/coder.c (fiasco_coder :85, alloc_coder :190, the name
 *  templates :390, video_coder :490, frame_coder :692) is organised here as a small job
 *  scheduler:
 *
 *    unit    = one frame of one stream (a stream is the whole picture, or one tile of it when
 *	        tile-split mode is asked for, see below);
 *    wave    = all units whose references exist already.  Intra frames have none, so all of them
 *	        are wave 0; a predicted frame follows the frames it is predicted from.  Every wave
 *	        is ONE kernel launch per GPU and frame type (one thread block per unit);
 *    GPUs    = the units of a wave are dealt round-robin to the devices in use, one host thread
 *	        per device; the finished automata land in host memory of this process, which is all
 *	        the "gather" a single process needs.
 *
 *  Everything below frame_coder's subdivide() call runs on the GPU through fb200_encode_tiles() /
 *  fb200_encode_predicted() (include/fiasco_b200.h); the automata are serialised by fco_writer.c.
 *
 *  Two opt-ins through the environment (the default stays one monolithic
 *  stream on device 0, byte-identical to the reference coder):
 *    FIASCO_GPUS=<n>	      use the first n CUDA devices (default 1);
 *    FIASCO_TILE_SPLIT=<k>   cut every frame into 2^k equal tiles, each coded as its own FIASCO
 *			      stream (exactly what the reference produces for the cropped pictures)
 *			      and written to <output>.tNN.<ext>, NN = row-major tile number.
 *
 *  Supported: stills and sequences, grey and colour (4:4:4), intra, P and B frames (full-pixel
 *  vectors), nondeterministic prediction of intra frames, the built-in initial basis "small.fco",
 *  rle domain pool, adaptive coefficient model, optimisation levels 0..2 of the command line.
 *  Everything else is refused with an error message (never silently approximated).
 */
#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <time.h>
#include <stdlib.h>
#include <string.h>

#include "fi_internal.h"
#include "fiasco_host.h"

#define fi_min(a, b) ((a) > (b) ? (b) : (a))
#define fi_max(a, b) ((a) < (b) ? (b) : (a))

#define FI_MAXGPUS 16

/*****************************************************************************
			      names of the input frames
*****************************************************************************/

typedef struct name_list
{
   char	  **v;
   unsigned n, cap;
} name_list_t;

static void
names_add (name_list_t *l, const char *prefix, size_t prefix_len, int number, int digits,
	   const char *suffix)
{
   char *s;

   if (l->n == l->cap)
   {
      l->cap = l->cap ? 2 * l->cap : 16;
      l->v   = realloc (l->v, l->cap * sizeof (char *));
      if (!l->v)
	 fi_error ("Out of memory!");
   }
   s = fiasco_calloc (prefix_len + strlen (suffix) + 24, 1);
   if (digits < 0)
      sprintf (s, "%.*s%s", (int) prefix_len, prefix, suffix);
   else
      sprintf (s, "%.*s%0*d%s", (int) prefix_len, prefix, digits, number, suffix);
   l->v [l->n++] = s;
}

static void
names_free (name_list_t *l)
{
   for (unsigned i = 0; i < l->n; i++)
      free (l->v [i]);
   free (l->v);
   memset (l, 0, sizeof *l);
}

/* a run of decimal digits at *s: its value, the number of digits; *s moves behind it */
static int
scan_number (const char **s, int *value)
{
   int	digits = 0;
   long v      = 0;

   while (isdigit ((unsigned char) **s))
   {
      v = v * 10 + (**s - '0');
      if (v > 100000000)
	 return 0;
      (*s)++;
      digits++;
   }
   *value = (int) v;
   return digits;
}

/*
 *  The array of names / templates the caller passed (doc/fiasco_coder.3: every element is a file
 *  name or "prefix[start-end{+,-}step]suffix", e.g. "img0[12-01-2].pgm" = img012.pgm, img010.pgm,
 *  ..., img002.pgm) as the flat list of frame names, in coding input order.  The numbers keep
 *  the width of 'start' (leading zeros).
 */
static void
expand_names (char const *const *templates, name_list_t *out)
{
   static const char *bad = "Input name template conversion failure.\n"
			    "Check spelling of template.";

   for (; *templates; templates++)
   {
      const char *t	= *templates;
      const char *open	= strchr (t, '[');
      const char *s;
      int	  start, end, step = 1, digits;

      if (!open)
      {
	 names_add (out, t, strlen (t), 0, -1, "");
	 continue;
      }
      s	     = open + 1;
      digits = scan_number (&s, &start);
      if (!digits || *s++ != '-' || !scan_number (&s, &end))
	 fi_error (bad);
      if (*s == '+' || *s == '-')
      {
	 const int down = *s++ == '-';

	 /* a step of 0 would never leave the first frame (the reference divides by it) */
	 if (!scan_number (&s, &step) || step == 0)
	    fi_error (bad);
	 if (down)
	    step = -step;
      }
      if (*s++ != ']' || (step > 0 ? start > end : start < end))
	 fi_error (bad);
      for (int v = start; step > 0 ? v <= end : v >= end; v += step)
	 names_add (out, t, (size_t) (open - t), v, digits, s);
   }
}

/*****************************************************************************
			  frame types and coding order
*****************************************************************************/

enum {T_INTRA = 0, T_P = 1, T_B = 2,
      T_ND = 3,		/* (kinds of workspace only: intra frames with nondeterministic prediction, */
      T_INTRA_SEQ = 4};	/* intra frames of a colour sequence that has predicted frames) */

static int
pattern_type (unsigned frame, const char *pattern)
{
   const int t = toupper ((unsigned char) pattern [frame % strlen (pattern)]);

   return frame == 0 || t == 'I' ? T_INTRA : t == 'P' ? T_P : t == 'B' ? T_B : -1;
}

/*
 *  The schedule of a sequence (video_coder, codec/coder.c:490-680).  B frames are held back until
 *  the next frame that is not a B frame -- their future reference -- has been coded; B frames
 *  that nothing follows lose their last member, which is coded as a P frame instead.  Then, in
 *  coding order, the references: a P frame is predicted from the frame coded just before it; a B
 *  frame keeps the references of its neighbourhood, where the frame coded last replaces the
 *  FUTURE one if it ran ahead of the display order, else (when B frames may serve as references)
 *  the PAST one (coder.c:571-627).
 *
 *  order [k]	 display number of the k-th coded frame
 *  ctype [n]	 type frame n is coded with
 *  past [n], future [n]   display numbers of its references or -1
 *  wave [n]	 length of the longest chain of references behind frame n
 */
typedef struct schedule
{
   unsigned *order;
   int	    *ctype, *past, *future, *wave;
   unsigned  n_waves;
} schedule_t;

static void
schedule_free (schedule_t *sc)
{
   free (sc->order);
   free (sc->ctype);
   free (sc->past);
   free (sc->future);
   free (sc->wave);
   memset (sc, 0, sizeof *sc);
}

static void
make_schedule (schedule_t *sc, unsigned frames, const char *pattern, int B_as_past_ref,
	       int chain_all)
{
   unsigned *held = fiasco_calloc (frames, sizeof (unsigned));
   unsigned  n_held = 0, coded = 0;
   int	     cur = -1, past = -1, future = -1, ahead = 0;
   unsigned  next_display = 0;
   uint8_t  *done = fiasco_calloc (frames + 1, 1);

   sc->order  = fiasco_calloc (frames, sizeof (unsigned));
   sc->ctype  = fiasco_calloc (frames, sizeof (int));
   sc->past   = fiasco_calloc (frames, sizeof (int));
   sc->future = fiasco_calloc (frames, sizeof (int));
   sc->wave   = fiasco_calloc (frames, sizeof (int));
   for (unsigned n = 0; n < frames; n++)
   {
      const int type = pattern_type (n, pattern);

      if (type == T_B && n + 1 < frames)
      {
	 held [n_held++] = n;
	 continue;
      }
      /* the frame all held B frames wait for (a trailing B frame becomes that P frame) */
      sc->ctype [n]	   = type == T_B ? T_P : type;
      sc->order [coded++] = n;
      for (unsigned i = 0; i < n_held; i++)
      {
	 sc->ctype [held [i]]  = T_B;
	 sc->order [coded++] = held [i];
      }
      n_held = 0;
   }
   for (unsigned k = 0; k < frames; k++)
   {
      const unsigned n = sc->order [k];

      if (sc->ctype [n] == T_P)
      {
	 past	= cur;
	 future = -1;
      }
      else if (sc->ctype [n] == T_B)
      {
	 if (ahead)
	    future = cur;
	 else if (B_as_past_ref)
	    past = cur;
      }
      else
	 past = future = -1;
      sc->past [n]   = sc->ctype [n] ? past : -1;
      sc->future [n] = sc->ctype [n] == T_B ? future : -1;
      done [n] = 1;
      ahead    = n > next_display;
      while (next_display < frames && done [next_display])
	 next_display++;
      cur = (int) n;
      /* position in the chain of references */
      sc->wave [n] = 0;
      if (sc->past [n] >= 0)
	 sc->wave [n] = sc->wave [sc->past [n]] + 1;
      if (sc->future [n] >= 0 && sc->wave [sc->future [n]] + 1 > sc->wave [n])
	 sc->wave [n] = sc->wave [sc->future [n]] + 1;
      if (chain_all && k)
	 sc->wave [n] = sc->wave [sc->order [k - 1]] + 1;
      if ((unsigned) sc->wave [n] + 1 > sc->n_waves)
	 sc->n_waves = (unsigned) sc->wave [n] + 1;
   }
   free (held);
   free (done);
}

/*****************************************************************************
				units, GPUs, waves
*****************************************************************************/

typedef struct unit		/* one frame of one stream */
{
   const int16_t *plane [3];
   fb200_wfa_t	  wfa;
   int16_t	 *recon;	/* the regenerated frame: reference of the frames predicted from it */
   uint8_t	 *delta;	/* delta flags of the states (predicted frames) */
   int		  is_reference;
} unit_t;

typedef struct gpu
{
   int		device;
   fb200_ctx_t *ctx [5];	/* workspaces for intra, P and B frames, intra frames with ND prediction */
   int		ctx_tiles [5];
   pthread_t	thread;
   int		failed;
   char		err [600];
} gpu_t;

typedef struct wave_work	/* what the GPU threads of one launch share */
{
   int			 type;		/* frame type: are there reference frames */
   int			 kind;		/* workspace: the frame type, or T_ND for intra frames of a job with
					   nondeterministic prediction */
   unsigned		 cnt, n_gpus;
   unit_t	       **u;
   unit_t	       **past, **future;
   const fb200_params_t *p;
   const fb200_motion_t *mo;
   unsigned		 width, height, bands;
   unsigned		 max_share;	/* units per GPU this job will ever put into one launch */
} wave_work_t;

typedef struct worker_arg
{
   gpu_t	     *g;
   unsigned	      index;
   const wave_work_t *w;
} worker_arg_t;

/* the units i = index, index + n_gpus, ... of the wave on this thread's device */
static void *
wave_worker (void *arg)
{
   const worker_arg_t *a = arg;
   const wave_work_t  *w = a->w;
   gpu_t	      *g = a->g;
   const unsigned      share = (w->cnt + w->n_gpus - 1 - a->index) / w->n_gpus;
   const int16_t     **planes, **past = NULL, **future = NULL;
   fb200_wfa_t	      *batch;
   unsigned	       i, k;
   int		       rc = FB200_OK;
   jmp_buf	       caller;	/* the helpers below are public entry points with a try of their own */

   g->failed = 0;
   if (!share)
      return NULL;
   planes = calloc ((size_t) share * w->bands, sizeof *planes);
   batch  = calloc (share, sizeof *batch);
   if (w->type)
   {
      past   = calloc (share, sizeof *past);
      future = calloc (share, sizeof *future);
   }
   if (!planes || !batch || (w->type && (!past || !future)))
   {
      snprintf (g->err, sizeof g->err, "Out of memory!");
      g->failed = 1;
      goto out;
   }
   for (i = a->index, k = 0; i < w->cnt; i += w->n_gpus, k++)
   {
      for (unsigned b = 0; b < w->bands; b++)
	 planes [k * w->bands + b] = w->u [i]->plane [b];
      batch [k] = w->u [i]->wfa;
      if (w->type)
      {
	 past [k]   = w->past [i]->recon;
	 future [k] = w->future [i] ? w->future [i]->recon : NULL;
      }
   }
   struct timespec t0, t1, t2;

   clock_gettime (CLOCK_MONOTONIC, &t0);
   if (!g->ctx [w->kind] || g->ctx_tiles [w->kind] < (int) share)
   {
      const int tiles = (int) fi_max (share, w->max_share);
      fb200_motion_t mo;

      if (g->ctx [w->kind])
	 fb200_destroy (g->ctx [w->kind]);
      g->ctx [w->kind] = NULL;
      if (w->kind)
      {
	 mo	       = *w->mo;
	 mo.frame_type = w->kind;
	 rc = fb200_create_predicted (&g->ctx [w->kind], w->p, &mo, tiles, g->device, g->err,
				      sizeof g->err);
      }
      else
	 rc = fb200_create (&g->ctx [w->kind], w->p, tiles, g->device, g->err, sizeof g->err);
      g->ctx_tiles [w->kind] = tiles;
   }
   clock_gettime (CLOCK_MONOTONIC, &t1);
   if (rc == FB200_OK)
   {
      if (w->kind)
	 rc = fb200_encode_predicted (g->ctx [w->kind], (int) share, planes, w->type ? past : NULL,
				      w->type == T_B ? future : NULL, batch, g->err, sizeof g->err);
      else
	 rc = fb200_encode_tiles (g->ctx [w->kind], (int) share, planes, batch, NULL, 0, NULL,
				  g->err, sizeof g->err);
   }
   clock_gettime (CLOCK_MONOTONIC, &t2);
   if (getenv ("FIASCO_TIMINGS"))
   {
      fb200_stats_t st;

      memset (&st, 0, sizeof st);
      if (g->ctx [w->kind])
	 fb200_get_stats (g->ctx [w->kind], &st);
      fprintf (stderr, "fiasco_coder:   gpu %d, %u units: context %.1f ms, call %.1f ms (h2d %.1f, kernel %.1f, "
	       "d2h %.1f)\n", g->device, share,
	       1e3 * (double) (t1.tv_sec - t0.tv_sec) + 1e-6 * (double) (t1.tv_nsec - t0.tv_nsec),
	       1e3 * (double) (t2.tv_sec - t1.tv_sec) + 1e-6 * (double) (t2.tv_nsec - t1.tv_nsec),
	       (double) st.h2d_ms, (double) st.kernel_ms, (double) st.d2h_ms);
   }
   if (rc != FB200_OK)
   {
      if (!g->err [0])
	 snprintf (g->err, sizeof g->err, "GPU encoder failed");
      g->failed = 1;
      goto out;
   }
   /* host part of the frame: close the holes of a predicted frame, derive its delta flags, and
      regenerate the picture if later frames are predicted from it (coder.c:642-651) */
   for (i = a->index, k = 0; i < w->cnt && !g->failed; i += w->n_gpus, k++)
   {
      unit_t		   *u = w->u [i];
      fiasco_frame_motion_t fm;

      u->wfa = batch [k];
      memset (&fm, 0, sizeof fm);
      memcpy (caller, fi_env, sizeof caller);
      if (w->kind)
      {
	 int ok = 0;

	 u->delta = calloc (FI_MAXSTATES, 1);
	 if (u->delta)
	    ok = fiasco_finish_predicted_frame (&u->wfa, u->wfa.mv_type, u->wfa.mv_fx, u->wfa.mv_fy,
						w->type == T_B ? u->wfa.mv_bx : NULL,
						w->type == T_B ? u->wfa.mv_by : NULL, u->delta);
	 memcpy (fi_env, caller, sizeof caller);
	 if (!ok)
	 {
	    snprintf (g->err, sizeof g->err, "%s",
		      u->delta ? fiasco_get_error_message () : "Out of memory!");
	    g->failed = 1;
	    break;
	 }
	 fm.frame_type	= w->type;
	 fm.mv_type	= u->wfa.mv_type;
	 fm.mv_fx	= u->wfa.mv_fx;
	 fm.mv_fy	= u->wfa.mv_fy;
	 fm.mv_bx	= u->wfa.mv_bx;
	 fm.mv_by	= u->wfa.mv_by;
	 fm.delta_state = u->delta;
      }
      if (u->is_reference)
      {
	 int ok = 0;

	 u->recon = calloc ((size_t) w->width * w->height * w->bands, sizeof (int16_t));
	 if (u->recon)
	    ok = (w->bands == 3 ? fiasco_regenerate_colour_frame : fiasco_regenerate_frame)
		    (&u->wfa, &fm, (int) w->width, (int) w->height, w->type ? past [k] : NULL,
		     w->type ? future [k] : NULL, u->recon);
	 memcpy (fi_env, caller, sizeof caller);
	 if (!ok)
	 {
	    snprintf (g->err, sizeof g->err, "%s",
		      u->recon ? fiasco_get_error_message () : "Out of memory!");
	    g->failed = 1;
	 }
      }
   }
out:
   free (planes);
   free (batch);
   free (past);
   free (future);
   return NULL;
}

/* everything fiasco_coder() holds while it runs; released in one place, on success and on
   failure alike (the library is not re-entrant, like the reference: lib/error.c:48-53) */
static struct
{
   fiasco_c_options_t *default_options;
   name_list_t	       names;
   fi_image_t	     **images;
   unsigned	       n_images;
   int16_t	     **crops;
   size_t	       n_crops;
   unit_t	      *units;
   size_t	       n_units;
   schedule_t	       sched;
   fi_bits_t	      *output;
   gpu_t	       gpus [FI_MAXGPUS];
   unsigned	       n_gpus;
   unit_t	     **wave_u, **wave_past, **wave_future;
   char		      *tile_name;
   fb200_params_t      prep_p;		/* workspaces set up beside the reading of the frames */
   unsigned	       prep_tiles;
   pthread_t	       prep_thread;
   int		       prep_running;
   pthread_t	       release_thread;
   int		       release_running;
   fi_bits_t	     **frame_bits;	/* the frames' own bit streams until they are joined */
   size_t	       n_frame_bits;
} job;

/* host thread: create the intra contexts of all GPUs of the job (errors are left to the launch,
   which creates what it does not find) */
static void *
prepare_contexts (void *arg)
{
   (void) arg;
   for (unsigned g = 0; g < job.n_gpus; g++)
   {
      gpu_t *gp = &job.gpus [g];
      char   err [600];

      if (!gp->ctx [0]
	  && fb200_create (&gp->ctx [0], &job.prep_p, (int) job.prep_tiles, gp->device, err, sizeof err) == FB200_OK)
	 gp->ctx_tiles [0] = (int) job.prep_tiles;
      else
	 gp->ctx [0] = NULL;
   }
   return NULL;
}

static void *
release_contexts (void *arg)
{
   (void) arg;
   for (unsigned g = 0; g < FI_MAXGPUS; g++)
      for (int t = 0; t < 5; t++)
	 if (job.gpus [g].ctx [t])
	 {
	    fb200_destroy (job.gpus [g].ctx [t]);
	    job.gpus [g].ctx [t] = NULL;
	 }
   return NULL;
}

static void
job_release (void)
{
   if (job.prep_running)
      pthread_join (job.prep_thread, NULL);
   if (job.release_running)
      pthread_join (job.release_thread, NULL);
   job.prep_running = job.release_running = 0;
   for (unsigned g = 0; g < FI_MAXGPUS; g++)
      for (int t = 0; t < 5; t++)
	 if (job.gpus [g].ctx [t])
	    fb200_destroy (job.gpus [g].ctx [t]);
   for (size_t i = 0; i < job.n_units; i++)
   {
      fb200_wfa_free (&job.units [i].wfa);
      free (job.units [i].recon);
      free (job.units [i].delta);
   }
   free (job.units);
   for (size_t i = 0; job.crops && i < job.n_crops; i++)
      free (job.crops [i]);
   free (job.crops);
   for (unsigned i = 0; i < job.n_images; i++)
      if (job.images [i])
	 fi_free_image (job.images [i]);
   free (job.images);
   names_free (&job.names);
   schedule_free (&job.sched);
   if (job.output)
      fi_bits_close (job.output);
   if (job.default_options)
      fiasco_c_options_delete (job.default_options);
   for (size_t i = 0; i < job.n_frame_bits; i++)
      fi_bits_free_mem (job.frame_bits [i]);
   free (job.frame_bits);
   free (job.wave_u);
   free (job.wave_past);
   free (job.wave_future);
   free (job.tile_name);
   memset (&job, 0, sizeof job);
}

/* run one launch on all GPUs of the job; the first failure becomes the error of the call */
static void
run_wave (wave_work_t *w)
{
   worker_arg_t arg [FI_MAXGPUS];
   unsigned	g, started = 0;

   w->n_gpus = job.n_gpus;
   for (g = 0; g < job.n_gpus; g++)
   {
      arg [g].g	    = &job.gpus [g];
      arg [g].index = g;
      arg [g].w	    = w;
   }
   /* one device: no thread at all (the common case, and the reference's own threading model) */
   if (job.n_gpus == 1)
      wave_worker (&arg [0]);
   else
   {
      for (g = 0; g < job.n_gpus; g++, started++)
	 if (pthread_create (&job.gpus [g].thread, NULL, wave_worker, &arg [g]))
	    break;
      for (g = 0; g < started; g++)
	 pthread_join (job.gpus [g].thread, NULL);
      if (started < job.n_gpus)
	 fi_error ("Can't start a host thread for GPU %u.", started);
   }
   for (g = 0; g < job.n_gpus; g++)
      if (job.gpus [g].failed)
	 fi_error ("%s", job.gpus [g].err);
}

/* the progress meter of subdivide() (subdivide.c:323-349) for one finished band: the bar always
   ends up as 50 marks; the percent counter shows the values the device recorded */
static void
draw_progress (fiasco_progress_e meter, const fb200_wfa_t *wfa, unsigned band)
{
   if (meter == FIASCO_PROGRESS_BAR)
      for (int i = 0; i < 50; i++)
	 fi_info ("#");
   else if (meter == FIASCO_PROGRESS_PERCENT)
      for (unsigned p = 1; p <= 100; p++)
	 if (wfa->progress [band][p >> 5] & (1u << (p & 31)))
	    fi_info ("%3d%%  \r", p);
   if (meter != FIASCO_PROGRESS_NONE)
      fi_message ("");
}

/* <output>.tNN.<ext> */
static char *
tile_output_name (const char *outputname, unsigned tile, unsigned tiles)
{
   const char *dot   = strrchr (outputname, '.');
   const char *slash = strrchr (outputname, '/');
   char	      *s     = fiasco_calloc (strlen (outputname) + 16, 1);
   const int   width = tiles > 100 ? 3 : 2;

   if (dot && (!slash || dot > slash))
      sprintf (s, "%.*s.t%0*u%s", (int) (dot - outputname), outputname, width, tile, dot);
   else
      sprintf (s, "%s.t%0*u", outputname, width, tile);
   return s;
}

/* FIASCO_TIMINGS=1: wall time of the phases of a call on stderr (measurement aid) */
static double
phase_clock (const char *what)
{
   static double   last;
   struct timespec ts;
   double	   now;

   clock_gettime (CLOCK_MONOTONIC, &ts);
   now = (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
   if (what && getenv ("FIASCO_TIMINGS"))
      fprintf (stderr, "fiasco_coder: %-28s %8.1f ms\n", what, 1e3 * (now - last));
   last = now;
   return now;
}

static unsigned
env_unsigned (const char *name, unsigned dflt, unsigned max)
{
   const char *e = getenv (name);
   char	      *end;
   long	       v;

   if (!e || !*e)
      return dflt;
   v = strtol (e, &end, 10);
   if (*end || v < 0 || v > (long) max)
      fi_error ("Environment variable %s=`%s': a number in 0..%u is expected.", name, e, max);
   return (unsigned) v;
}

/* frame n of the job: read it, cut it into the streams' pictures (host threads, side by side) */
typedef struct read_ctx
{
   unsigned split, streams, frames, bands, cols, tw, th, width;
} read_ctx_t;

static void
read_frame (size_t n, void *ctx)
{
   const read_ctx_t *r = ctx;

   job.images [n] = fi_read_image (job.names.v [n]);
   for (unsigned s = 0; s < r->streams; s++)
   {
      unit_t *u = &job.units [(size_t) s * r->frames + n];

      if (fb200_wfa_alloc (&u->wfa, FI_MAXSTATES))
	 fi_error ("Out of memory!");
      for (unsigned b = 0; b < r->bands; b++)
	 if (!r->split)
	    u->plane [b] = job.images [n]->pixels [b];
	 else
	 {
	    const unsigned x0 = (s % r->cols) * r->tw, y0 = (s / r->cols) * r->th;
	    int16_t	  *c  = fiasco_calloc ((size_t) r->tw * r->th, sizeof (int16_t));

	    job.crops [((size_t) n * r->streams + s) * r->bands + b] = c;
	    for (unsigned y = 0; y < r->th; y++)
	       memcpy (c + (size_t) y * r->tw,
		       job.images [n]->pixels [b] + (size_t) (y0 + y) * r->width + x0,
		       (size_t) r->tw * sizeof (int16_t));
	    u->plane [b] = c;
	 }
   }
   if (r->split)		/* the tiles hold their own copies */
   {
      fi_free_image (job.images [n]);
      job.images [n] = NULL;
   }
}

/* frame 'coded' (in coding order) of stream s as a bit stream of its own (host threads) */
typedef struct write_ctx
{
   unsigned	       frames, bands;
   const fi_wfainfo_t *wi;
   const c_options_t  *cop;
   fi_bits_t	     **out;	/* [streams * frames] */
   unsigned	       phase;	/* position of the frame's first bit in its byte of the file */
} write_ctx_t;

static void
write_frame (size_t i, void *ctx)
{
   const write_ctx_t *wc    = ctx;
   const unsigned     s	    = (unsigned) (i / wc->frames), coded = (unsigned) (i % wc->frames);
   const unsigned     n	    = (unsigned) job.sched.order [coded];
   unit_t	     *u	    = &job.units [(size_t) s * wc->frames + n];
   fi_wfa_t	      w;

   memset (&w, 0, sizeof w);	/* intra frame: no motion data */
   if (job.sched.ctype [n])
   {
      w.frame_type  = job.sched.ctype [n];
      w.mv_bx	    = (const int8_t (*)[2]) u->wfa.mv_bx;
      w.mv_by	    = (const int8_t (*)[2]) u->wfa.mv_by;
      w.x	    = (const uint16_t (*)[2]) u->wfa.x;
      w.y	    = (const uint16_t (*)[2]) u->wfa.y;
      w.mv_type	    = (const int8_t (*)[2]) u->wfa.mv_type;
      w.mv_fx	    = (const int8_t (*)[2]) u->wfa.mv_fx;
      w.mv_fy	    = (const int8_t (*)[2]) u->wfa.mv_fy;
      w.delta_state = u->delta;
   }
   else if (u->delta)		/* an intra frame with nondeterministic prediction */
      w.delta_state = u->delta;
   w.info	    = wc->wi;
   w.states	    = u->wfa.states;
   w.basis_states   = u->wfa.basis_states;
   w.root_state	    = u->wfa.root_state;
   w.level_of_state = u->wfa.level_of_state;
   w.domain_type    = u->wfa.domain_type;
   w.tree	    = (const int16_t (*)[2]) u->wfa.tree;
   w.into	    = (const int16_t (*)[2][6]) u->wfa.into;
   w.weight	    = (const float (*)[2][6]) u->wfa.weight;
   w.y_state	    = (const int16_t (*)[2]) u->wfa.y_state;
   w.y_column	    = (const uint8_t (*)[2]) u->wfa.y_column;
   /* (the stream's byte alignments count from the start of the file: 'phase' = bits of the frame's
      first byte that belong to the frame before it) */
   wc->out [i]	    = fi_bits_open_mem (wc->phase);
   fi_write_next_wfa (&w, n, coded == 0, wc->cop->normal_domains, wc->cop->delta_domains, wc->out [i]);
}

static int
coder (char const *const *inputname, const char *outputname, float quality,
       const fiasco_c_options_t *options)
{
   char const *const  default_input [] = {"-", NULL};
   char const *const *templates;
   const c_options_t *cop;
   fi_wfainfo_t	      wi;
   fb200_params_t     p;
   fb200_motion_t     mo;
   unsigned	      frames, width = 0, height = 0, n, bands, n_predicted = 0;
   unsigned	      split, gpus, streams, cols, rows, tw, th;
   int		      color = 0, nd = 0;

   if (!inputname || !inputname [0] || strcmp (inputname [0], "-") == 0)
      templates = default_input;
   else
      templates = inputname;
   if (quality <= 0)
   {
      fi_set_error ("Compression quality has to be positive.");
      return 0;
   }
   else if (quality >= 100)
      fi_warning ("Quality typically is 1 (worst) to 100 (best).\n"
		  "Be prepared for a long running time.");
   if (options)
   {
      cop = fi_cast_c_options (options);
      if (!cop)
	 return 0;
   }
   else
   {
      job.default_options = fiasco_c_options_new ();
      cop		  = fi_cast_c_options (job.default_options);
   }
   split = env_unsigned ("FIASCO_TILE_SPLIT", 0, 12);
   gpus	 = env_unsigned ("FIASCO_GPUS", 1, FI_MAXGPUS);
   if (!split)
   {
      job.output = fi_bits_open (outputname);
      if (!job.output)
      {
	 fi_set_error ("Can't write outputfile `%s'.\n%s",
		       outputname ? outputname : "<stdout>", fi_system_error ());
	 return 0;
      }
   }
   else if (!outputname || strcmp (outputname, "-") == 0)
      fi_error ("FIASCO_TILE_SPLIT writes one file per tile: an output file name is needed.");

   phase_clock (NULL);
   /* all frames readable, same size, same colour model (coder.c:204-240) */
   expand_names (templates, &job.names);
   frames = job.names.n;
   if (!frames)
      fi_error ("No input frames.");
   for (n = 0; n < frames; n++)
   {
      unsigned w, h;
      int      c;

      fi_read_pnm_header (job.names.v [n], &w, &h, &c);
      if (n)
      {
	 if (w != width || h != height)
	    fi_error ("`%s': all images of a sequence have to be of the same size.",
		      job.names.v [n]);
	 if (c != color)
	    fi_error ("`%s': all images a sequence have to use the same color model.",
		      job.names.v [n]);
      }
      else
      {
	 width	= w;
	 height = h;
	 color	= c;
      }
   }
   bands = color ? 3 : 1;

   /* what this build does not do is refused, not approximated */
   for (n = 0; n < frames; n++)
      if (pattern_type (n, cop->pattern) < 0)
	 fi_error ("Frame type %c not valid. Choose one of I,B or P.",
		   cop->pattern [n % strlen (cop->pattern)]);
   /* the frames of a colour sequence are chained through the coder's options: the chroma bands
      raise lc_min_level for good (coder.c:797), so frame k starts where frame k - 1 ended */
   make_schedule (&job.sched, frames, cop->pattern, cop->B_as_past_ref, color && frames > 1);
   for (n = 0; n < frames; n++)
      if (job.sched.ctype [n])
      {
	 if (cop->half_pixel_prediction)
	    fi_error ("Half pixel motion compensation is not available in the B200 build.");
	 if (!cop->normal_domains || !cop->delta_domains
	     || cop->d_rpf_mantissa != cop->rpf_mantissa || cop->d_rpf_range != cop->rpf_range
	     || cop->d_dc_rpf_mantissa != cop->dc_rpf_mantissa
	     || cop->d_dc_rpf_range != cop->dc_rpf_range)
	    fi_error ("Predicted frames: only the default domain pool and quantisation "
		      "settings of the prediction errors are available.");
	 /* e.g. a B frame whose future reference is an I frame: the reference coder drops the
	    past frame there (coder.c:581-591) and then reads through the NULL pointer */
	 if (job.sched.past [n] < 0 || (job.sched.ctype [n] == T_B && job.sched.future [n] < 0))
	    fi_error ("Frame %d (pattern `%s') has no reference frame to be predicted from.",
		      n, cop->pattern);
	 n_predicted++;
      }
   /* `--prediction' (nd_prediction, prediction.c:371) is tried on the intra frames of grey input
      (coder.c:741-743, 806-807: never on a colour frame, whose delta pool merely changes its kind
      -- no difference while the dictionary cannot fill up) */
   nd = cop->prediction && !color;
   if (cop->prediction && color && cop->max_states < FI_MAXSTATES)
      fi_error ("`--prediction' on colour frames is available with the full dictionary size only.");
   if (nd && (!cop->normal_domains || !cop->delta_domains
	      || cop->d_rpf_mantissa != cop->rpf_mantissa || cop->d_rpf_range != cop->rpf_range
	      || cop->d_dc_rpf_mantissa != cop->dc_rpf_mantissa
	      || cop->d_dc_rpf_range != cop->dc_rpf_range))
      fi_error ("Nondeterministic prediction: only the default domain pool and quantisation "
		"settings of the prediction errors are available.");
   if (cop->full_search)
      fi_error ("Optimization level 3 (full search) is not available: the reference "
		"coder's behaviour is undefined there.");
   if (strcmp (cop->basis_name, "small.fco") != 0)
      fi_error ("Initial basis `%s' is not available, only the built-in `small.fco'.",
		cop->basis_name);
   if (strcasecmp (cop->id_domain_pool, "rle") != 0
       || strcasecmp (cop->id_rpf_model, "adaptive") != 0)
      fi_error ("Only the `rle' domain pool and the `adaptive' coefficients model are "
		"available.");

   /* tile-split mode: 2^split equal tiles, each a picture the reference would accept */
   cols	   = 1u << ((split + 1) / 2);
   rows	   = 1u << (split / 2);
   streams = cols * rows;
   if (width % cols || height % rows)
      fi_error ("FIASCO_TILE_SPLIT=%u: %ux%u pixels do not divide into %ux%u tiles.", split,
		width, height, cols, rows);
   tw = width / cols;
   th = height / rows;
   if (split && ((tw & 1) || (th & 1) || tw < 32 || th < 32))
      fi_error ("FIASCO_TILE_SPLIT=%u: tiles of %ux%u pixels are too small or not even.", split,
		tw, th);
   {
      const int have = fb200_device_count ();

      if (have < 1)
	 fi_error ("no CUDA device available (this library has no CPU path)");
      job.n_gpus = fi_min (gpus ? gpus : 1, (unsigned) have);
      for (n = 0; n < job.n_gpus; n++)
	 job.gpus [n].device = (int) n;
   }

   /* geometry and option clamping (coder.c:249-327), per stream */
   memset (&p, 0, sizeof p);
   memset (&wi, 0, sizeof wi);
   {
      unsigned lx = (unsigned) (log2 ((double) (tw - 1)) + 1);
      unsigned ly = (unsigned) (log2 ((double) (th - 1)) + 1);

      wi.level = fi_max (lx, ly) * 2 - ((ly == lx + 1) ? 1 : 0);
   }
   p.width	  = (int) tw;
   p.height	  = (int) th;
   p.bands	  = (int) bands;
   p.level	  = (int) wi.level;
   p.lc_min_level = (int) fi_max (cop->lc_min_level, 3);
   p.lc_max_level = (int) fi_min (cop->lc_max_level, wi.level - 1);
   /* the reference's tiling object never stores its exponent (tiling.c:68-91): the
      exponent is 0 whatever the caller set, so the tiling clamp is a no-op */
   if (p.lc_min_level > p.lc_max_level)
      p.lc_min_level = p.lc_max_level;
   wi.p_min_level = fi_max (cop->p_min_level, (unsigned) p.lc_min_level);
   wi.p_max_level = fi_min (cop->p_max_level, (unsigned) p.lc_max_level);
   if (wi.p_min_level > wi.p_max_level)
      wi.p_min_level = wi.p_max_level;
   p.images_level	= (int) fi_min (cop->images_level, (unsigned) p.lc_max_level - 1);
   wi.max_states	= fi_max (fi_min (cop->max_states, FI_MAXSTATES), 1);
   p.max_states		= (int) wi.max_states;
   p.max_elements	= (int) fi_max (fi_min (cop->max_elements, FI_MAXEDGES), 1);
   wi.chroma_max_states = fi_max (1, cop->chroma_max_states);
   p.chroma_max_states	= (int) wi.chroma_max_states;
   p.price		= 128 * 64 / quality;		/* coder.c:164 */
   p.chroma_decrease	= cop->chroma_decrease;
   wi.rpf      = fi_make_rpf (cop->rpf_mantissa, (int) cop->rpf_range);
   wi.dc_rpf   = fi_make_rpf (cop->dc_rpf_mantissa, (int) cop->dc_rpf_range);
   wi.d_rpf    = fi_make_rpf (cop->d_rpf_mantissa, (int) cop->d_rpf_range);
   wi.d_dc_rpf = fi_make_rpf (cop->d_dc_rpf_mantissa, (int) cop->d_dc_rpf_range);
   p.rpf_mantissa	 = (int) wi.rpf.mantissa_bits;
   p.rpf_range		 = wi.rpf.range;
   p.dc_rpf_mantissa	 = (int) wi.dc_rpf.mantissa_bits;
   p.dc_rpf_range	 = wi.dc_rpf.range;
   p.second_domain_block = cop->second_domain_block;
   p.state_capacity	 = 0;
   wi.basis_name    = cop->basis_name;
   wi.title	    = cop->title;
   wi.comment	    = cop->comment;
   wi.color	    = color;
   wi.width	    = tw;
   wi.height	    = th;
   wi.frames	    = frames;
   wi.fps	    = cop->fps;
   wi.search_range  = cop->search_range;
   wi.half_pixel    = cop->half_pixel_prediction;
   wi.B_as_past_ref = cop->B_as_past_ref;
   wi.smoothing	    = cop->smoothing;
   wi.nd_prediction = cop->prediction;
   mo.frame_type   = T_P;
   mo.p_min_level  = (int) wi.p_min_level;
   mo.p_max_level  = (int) wi.p_max_level;
   mo.search_range = (int) cop->search_range;

   phase_clock ("headers, options");
   /* the device workspaces of the intra frames are set up while the frames are read: every GPU
      gets the context of its largest intra launch (the launches find them, wave_worker) */
   {
      unsigned intra = 0;

      for (unsigned wv = 0; wv < job.sched.n_waves; wv++)
      {
	 unsigned cnt = 0;

	 for (n = 0; n < frames; n++)
	    if ((unsigned) job.sched.wave [n] == wv && job.sched.ctype [n] == 0)
	       cnt += streams;
	 intra = fi_max (intra, cnt);
      }
      job.prep_p     = p;
      job.prep_tiles = (intra + job.n_gpus - 1) / job.n_gpus;
      /* (FIASCO_PREPARE=1: measured slower on the bench box -- the allocations of the context
	 and the page faults of the reading threads contend for the address space) */
      if (job.prep_tiles && getenv ("FIASCO_PREPARE") && atoi (getenv ("FIASCO_PREPARE"))
	  && pthread_create (&job.prep_thread, NULL, prepare_contexts, NULL) == 0)
	 job.prep_running = 1;
   }
   /* read the frames, cut them into the streams' pictures */
   job.images	= fiasco_calloc (frames, sizeof (fi_image_t *));
   job.n_images = frames;
   job.units	= fiasco_calloc ((size_t) streams * frames, sizeof (unit_t));
   job.n_units	= (size_t) streams * frames;
   if (split)
      job.crops = fiasco_calloc ((size_t) streams * frames * bands, sizeof (int16_t *));
   {
      read_ctx_t rc = {split, streams, frames, bands, cols, tw, th, width};
      int	 serial = 0;

      if (split)
	 job.n_crops = (size_t) streams * frames * bands;
      for (n = 0; n < frames; n++)		/* standard input is read in order */
	 if (!job.names.v [n] || strcmp (job.names.v [n], "-") == 0)
	    serial = 1;
      if (serial)
	 for (n = 0; n < frames; n++)
	    read_frame (n, &rc);
      else if (fi_parallel_for (frames, read_frame, &rc))
	 fi_rethrow ();
   }
   for (n = 0; n < frames; n++)
      for (unsigned s = 0; s < streams; s++)
      {
	 if (job.sched.past [n] >= 0)
	    job.units [(size_t) s * frames + job.sched.past [n]].is_reference = 1;
	 if (job.sched.future [n] >= 0)
	    job.units [(size_t) s * frames + job.sched.future [n]].is_reference = 1;
      }

   if (job.prep_running)
   {
      pthread_join (job.prep_thread, NULL);
      job.prep_running = 0;
   }
   phase_clock ("frames read");
   /* the waves: every frame whose references are ready, over all streams, in one launch per
      frame type and GPU */
   job.wave_u	   = fiasco_calloc ((size_t) streams * frames, sizeof (unit_t *));
   job.wave_past   = fiasco_calloc ((size_t) streams * frames, sizeof (unit_t *));
   job.wave_future = fiasco_calloc ((size_t) streams * frames, sizeof (unit_t *));
   {
      /* the largest launch of every type sizes the workspaces once */
      unsigned max_cnt [3] = {0, 0, 0};

      for (unsigned wv = 0; wv < job.sched.n_waves; wv++)
      {
	 unsigned cnt [3] = {0, 0, 0};

	 for (n = 0; n < frames; n++)
	    if ((unsigned) job.sched.wave [n] == wv)
	       cnt [job.sched.ctype [n]] += streams;
	 for (int t = 0; t < 3; t++)
	    max_cnt [t] = fi_max (max_cnt [t], cnt [t]);
      }
      for (unsigned wv = 0; wv < job.sched.n_waves; wv++)
	 for (int type = 0; type < 3; type++)
	 {
	    wave_work_t w;

	    memset (&w, 0, sizeof w);
	    for (n = 0; n < frames; n++)
	    {
	       if ((unsigned) job.sched.wave [n] != wv || job.sched.ctype [n] != type)
		  continue;
	       for (unsigned s = 0; s < streams; s++)
	       {
		  unit_t *base = &job.units [(size_t) s * frames];

		  job.wave_u [w.cnt]	  = base + n;
		  job.wave_past [w.cnt]	  = type ? base + job.sched.past [n] : NULL;
		  job.wave_future [w.cnt] = type == T_B ? base + job.sched.future [n] : NULL;
		  /* a colour frame starts with the range levels its predecessor ended with */
		  if (color && wv)
		  {
		     const fb200_wfa_t *before = &base [job.sched.order [wv - 1]].wfa;

		     base [n].wfa.lc_min_level = before->lc_min_level;
		     /* ... and with what it left in the y_column entries of the state numbers
			(fb200_wfa_t.y_column_history) */
		     if (base [n].wfa.y_column_history && before->y_column_history)
			memcpy (base [n].wfa.y_column_history, before->y_column_history, 2 * FI_MAXSTATES);
		  }
		  w.cnt++;
	       }
	    }
	    if (!w.cnt)
	       continue;
	    w.type	= type;
	    w.kind	= type == T_INTRA && nd ? T_ND : type == T_INTRA && color && frames > 1 ? T_INTRA_SEQ : type;
	    w.u		= job.wave_u;
	    w.past	= job.wave_past;
	    w.future	= job.wave_future;
	    w.p		= &p;
	    w.mo	= &mo;
	    w.width	= tw;
	    w.height	= th;
	    w.bands	= bands;
	    w.max_share = (max_cnt [type] + job.n_gpus - 1) / job.n_gpus;
	    run_wave (&w);
	 }
   }
   phase_clock ("launches (contexts, copies)");
   /* the device memory is not needed any longer.  (Handing it back on a thread of its own beside
      the stream writer was measured: unpinning the staging buffers stalls the writer's threads, one
      call in three took 0.85 s instead of 0.09 s to write its streams.) */
   release_contexts (NULL);

   phase_clock ("contexts released");
   /* the streams: every frame is coded into a bit stream of its own on the host threads, then
      the frames of a stream are joined in coding order */
   fi_write_tables_init ();
   job.frame_bits   = fiasco_calloc ((size_t) streams * frames, sizeof (fi_bits_t *));
   job.n_frame_bits = (size_t) streams * frames;
   {
      write_ctx_t wc = {frames, bands, &wi, cop, job.frame_bits, 0};

      if (fi_parallel_for ((size_t) streams * frames, write_frame, &wc))
	 fi_rethrow ();
   }
   for (unsigned s = 0; s < streams; s++)
   {
      if (split)
      {
	 job.tile_name = tile_output_name (outputname, s, streams);
	 job.output    = fi_bits_open (job.tile_name);
	 if (!job.output)
	    fi_error ("Can't write outputfile `%s'.\n%s", job.tile_name, fi_system_error ());
      }
      for (unsigned coded = 0; coded < frames; coded++)
      {
	 unit_t *u;

	 n = job.sched.order [coded];
	 u = &job.units [(size_t) s * frames + n];
	 for (unsigned b = 0; b < bands; b++)
	    draw_progress (cop->progress_meter, &u->wfa, b);
	 fi_debug_message ("WFA contains %d states (%d basis states).", u->wfa.states,
			   u->wfa.basis_states);
	 fi_debug_message ("Total costs : %.2f", (double) u->wfa.costs [0]);
	 /* a frame without any edge ends with its matrices, not on a byte boundary: the frame after
	    it is coded again, with its byte alignments where they fall in the file */
	 if ((job.output->nbits & 7) != job.frame_bits [(size_t) s * frames + coded]->skip)
	 {
	    write_ctx_t again = {frames, bands, &wi, cop, job.frame_bits, (unsigned) (job.output->nbits & 7)};

	    fi_bits_free_mem (job.frame_bits [(size_t) s * frames + coded]);
	    job.frame_bits [(size_t) s * frames + coded] = NULL;
	    write_frame ((size_t) s * frames + coded, &again);
	 }
	 fi_bits_append (job.output, job.frame_bits [(size_t) s * frames + coded]);
	 fi_bits_free_mem (job.frame_bits [(size_t) s * frames + coded]);
	 job.frame_bits [(size_t) s * frames + coded] = NULL;
      }
      fi_bits_close (job.output);
      job.output = NULL;
      free (job.tile_name);
      job.tile_name = NULL;
   }
   if (job.release_running)
   {
      pthread_join (job.release_thread, NULL);
      job.release_running = 0;
   }
   phase_clock ("streams written");
   (void) n_predicted;
   return 1;
}

int
fiasco_coder (char const *const *inputname, const char *outputname, float quality,
	      const fiasco_c_options_t *options)
{
   int ok = 0;

   memset (&job, 0, sizeof job);
   fi_try
   {
      ok = coder (inputname, outputname, quality, options);
   }
   fi_catch
   {
      ok = 0;
   }
   job_release ();
   phase_clock ("everything released");
   return ok;
}

/*****************************************************************************
		      fiasco_host.h: stream writing for callers of the C ABI
*****************************************************************************/


void
fiasco_stream_info_init (fiasco_stream_info_t *info, const fb200_params_t *p)
{
   memset (info, 0, sizeof *info);
   info->width		   = p->width;
   info->height		   = p->height;
   info->color		   = p->bands == 3;
   info->max_states	   = (unsigned) p->max_states;
   info->chroma_max_states = (unsigned) p->chroma_max_states;
   /* the CLI passes prediction levels [6,10]; alloc_coder() clamps them into the range
      levels (coder.c:285-288) */
   info->p_min_level = fi_max (6u, (unsigned) p->lc_min_level);
   info->p_max_level = fi_min (10u, (unsigned) p->lc_max_level);
   if (info->p_min_level > info->p_max_level)
      info->p_min_level = info->p_max_level;
   info->smoothing	 = 70;
   info->fps		 = 25;
   info->rpf_mantissa	 = p->rpf_mantissa;
   info->rpf_range_e	 = p->rpf_range == 0.75f ? 0 : p->rpf_range == 1.5f ? 2
			   : p->rpf_range == 2.0f ? 3 : 1;
   info->dc_rpf_mantissa = p->dc_rpf_mantissa;
   info->dc_rpf_range_e	 = p->dc_rpf_range == 0.75f ? 0 : p->dc_rpf_range == 1.5f ? 2
			   : p->dc_rpf_range == 2.0f ? 3 : 1;
}

int
fiasco_write_stream (const char *filename, const fiasco_stream_info_t *info,
		     const fb200_wfa_t *frames, int n_frames)
{
   return fiasco_write_video_stream (filename, info, frames, NULL, n_frames, 16);
}

int
fiasco_write_video_stream (const char *filename, const fiasco_stream_info_t *info,
			   const fb200_wfa_t *frames, const fiasco_frame_motion_t *motion,
			   int n_frames, unsigned search_range)
{
   fi_bits_t *volatile out = NULL;

   fi_try
   {
      fi_wfainfo_t wi;
      int	   n;

      if (!info || !frames || n_frames < 1)
      {
	 fi_set_error ("fiasco_write_video_stream: bad arguments");
	 return 0;
      }
      memset (&wi, 0, sizeof wi);
      wi.basis_name	   = "small.fco";
      wi.title		   = info->title ? info->title : "";
      wi.comment	   = info->comment ? info->comment : "";
      wi.max_states	   = info->max_states;
      wi.chroma_max_states = info->chroma_max_states;
      wi.color		   = info->color;
      wi.width		   = (unsigned) info->width;
      wi.height		   = (unsigned) info->height;
      wi.rpf	  = fi_make_rpf ((unsigned) info->rpf_mantissa, info->rpf_range_e);
      wi.dc_rpf	  = fi_make_rpf ((unsigned) info->dc_rpf_mantissa, info->dc_rpf_range_e);
      wi.d_rpf	  = fi_make_rpf (3, 2);		/* options.c:86-89 defaults */
      wi.d_dc_rpf = fi_make_rpf (5, 1);
      wi.frames	       = (unsigned) n_frames;
      wi.fps	       = info->fps;
      wi.p_min_level   = info->p_min_level;
      wi.p_max_level   = info->p_max_level;
      wi.search_range  = search_range;
      wi.half_pixel    = 0;
      wi.B_as_past_ref = 1;
      wi.smoothing     = info->smoothing;
      wi.nd_prediction = info->nd_prediction;
      out = fi_bits_open (filename);
      if (!out)
      {
	 fi_set_error ("Can't write outputfile `%s'.\n%s", filename ? filename : "<stdout>",
		       fi_system_error ());
	 return 0;
      }
      for (n = 0; n < n_frames; n++)
      {
	 fi_wfa_t w;

	 if (frames [n].status != FB200_OK)
	    fi_error ("frame %d holds no automaton (status %d)", n, frames [n].status);
	 memset (&w, 0, sizeof w);
	 w.info		  = &wi;
	 w.states	  = frames [n].states;
	 w.basis_states	  = frames [n].basis_states;
	 w.root_state	  = frames [n].root_state;
	 w.level_of_state = frames [n].level_of_state;
	 w.domain_type	  = frames [n].domain_type;
	 w.tree		  = (const int16_t (*)[2]) frames [n].tree;
	 w.into		  = (const int16_t (*)[2][6]) frames [n].into;
	 w.weight	  = (const float (*)[2][6]) frames [n].weight;
	 w.y_state	  = (const int16_t (*)[2]) frames [n].y_state;
	 w.y_column	  = (const uint8_t (*)[2]) frames [n].y_column;
	 if (motion && motion [n].frame_type != 0)
	 {
	    if (motion [n].frame_type < 1 || motion [n].frame_type > 2)
	       fi_error ("frame %d: frame type %d", n, motion [n].frame_type);
	    w.frame_type  = motion [n].frame_type;
	    w.mv_bx	  = (const int8_t (*)[2]) motion [n].mv_bx;
	    w.mv_by	  = (const int8_t (*)[2]) motion [n].mv_by;
	    w.x		  = (const uint16_t (*)[2]) frames [n].x;
	    w.y		  = (const uint16_t (*)[2]) frames [n].y;
	    w.mv_type	  = (const int8_t (*)[2]) motion [n].mv_type;
	    w.mv_fx	  = (const int8_t (*)[2]) motion [n].mv_fx;
	    w.mv_fy	  = (const int8_t (*)[2]) motion [n].mv_fy;
	    w.delta_state = motion [n].delta_state;
	 }
	 else if (motion && info->nd_prediction)
	    w.delta_state = motion [n].delta_state;	/* differences of DC-predicted ranges */
	 fi_write_next_wfa (&w, motion ? (unsigned) motion [n].frame_number : (unsigned) n, n == 0, 1, 1, out);
      }
      fi_bits_close (out);
      return 1;
   }
   fi_catch
   {
      if (out)
	 fi_bits_close (out);
      return 0;
   }
}
