/*
 *  coder_api.c -- fiasco_coder(): the public entry point of the encoder.
 *
 *  Host-side control flow in the order of the reference (codec/coder.c:85-187 fiasco_coder,
 *  :190-372 alloc_coder, :390-488 input name templates, :490-690 video_coder, :692-892
 *  frame_coder); everything below frame_coder's subdivide() call runs on the GPU through
 *  fb200_encode_tiles() / fb200_encode_predicted() (include/fiasco_b200.h), the finished
 *  automaton comes back and is serialised by fco_writer.c.
 *
 *  Supported: still images and sequences, grey and colour (4:4:4) intra frames, P and B frames
 *  of grey sequences (full-pixel vectors), the built-in initial basis "small.fco", rle domain
 *  pool, adaptive coefficient model, optimisation levels 0..2 of the command line.  Everything
 *  else is refused with an error message (never silently approximated).
 */
#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fi_internal.h"
#include "fiasco_host.h"

#define fi_min(a, b) ((a) > (b) ? (b) : (a))
#define fi_max(a, b) ((a) < (b) ? (b) : (a))

/* "prefix[start-end{+,-}step]suffix" -> name of the i-th frame, or NULL */
static char *
input_name (char const *const *templptr, unsigned ith_image)
{
   const char *bad = "Input name template conversion failure.\nCheck spelling of template.";

   while (*templptr)
   {
      const char *template = *templptr++;
      const char *open	   = strchr (template, '[');

      if (!open)
      {
	 if (ith_image == 0)
	    return strdup (template);
	 ith_image--;
	 continue;
      }
      {
	 const char *s = open + 1, *s2;
	 unsigned    n_digits = 0;
	 int	     first, last, increment = 1, image_num;

	 for (s2 = s; isdigit ((unsigned char) *s2); s2++)
	    n_digits++;
	 if (sscanf (s, "%d", &first) != 1 || first < 0 || *s2++ != '-')
	    fi_error (bad);
	 s = s2;
	 while (isdigit ((unsigned char) *s2))
	    s2++;
	 if (sscanf (s, "%d", &last) != 1 || last < 0)
	    fi_error (bad);
	 if (*s2 == '+' || *s2 == '-')
	 {
	    s = s2++;
	    while (isdigit ((unsigned char) *s2))
	       s2++;
	    if (sscanf (s, "%d", &increment) != 1)
	       fi_error (bad);
	 }
	 if (*s2 != ']')
	    fi_error (bad);
	 image_num = first + increment * (int) ith_image;
	 if (image_num < 0)
	    fi_error (bad);
	 if ((increment > 0 && image_num > last) || (increment <= 0 && image_num < last))
	    ith_image -= (unsigned) ((last - first) / increment + 1);
	 else
	 {
	    char *name = fiasco_calloc (strlen (template) + 32, 1);

	    sprintf (name, "%.*s%0*d%s", (int) (open - template), template,
		     (int) n_digits, image_num, s2 + 1);
	    return name;
	 }
      }
   }
   return NULL;
}



/* 0 = I, 1 = P, 2 = B by the pattern (frame 0 is always intra) */
static int
pattern_type (unsigned frame, const char *pattern)
{
   const int t = toupper ((unsigned char) pattern [frame % strlen (pattern)]);

   return frame == 0 || t == 'I' ? 0 : t == 'P' ? 1 : t == 'B' ? 2 : -1;
}

/*
 *  The order in which video_coder() codes the frames (coder.c:490-680): display order, except
 *  that a B frame waits for the next non-B frame (its future reference), which is coded first;
 *  a B frame at the very end of the sequence is coded as a P frame.  order [k] = display number
 *  of the k-th coded frame, ctype [display number] = type it is coded with.
 */
static void
coding_order (unsigned frames, const char *pattern, unsigned *order, int *ctype)
{
   int display = 0, future_display = -1, coded = 0;

   while (display < (int) frames)
   {
      int type = pattern_type ((unsigned) display, pattern), frame;

      if (display == future_display)		/* already coded as a future reference */
      {
	 display++;
	 continue;
      }
      else if (type == 2 && display > future_display)
      {
	 int i = display;

	 frame = display;
	 while (type == 2)
	 {
	    i++;
	    if (i >= (int) frames)
	    {
	       future_display = i - 1;
	       type	      = 1;
	    }
	    else
	    {
	       future_display = i;
	       type	      = pattern_type ((unsigned) i, pattern);
	    }
	    frame = future_display;
	 }
      }
      else
      {
	 frame = display;
	 display++;
      }
      order [coded++] = (unsigned) frame;
      ctype [frame]   = type;
   }
}

int
fiasco_coder (char const *const *inputname, const char *outputname, float quality,
	      const fiasco_c_options_t *options)
{
   fi_try
   {
      char const *const	  default_input [] = {"-", NULL};
      char const *const  *template;
      fiasco_c_options_t *default_options = NULL;
      const c_options_t	 *cop;
      fi_bits_t		 *output;
      fi_wfainfo_t	  wi;
      fb200_params_t	  p;
      fb200_ctx_t	 *ctx = NULL;
      fb200_wfa_t	 *wfas;
      fi_image_t	**images;
      const int16_t	**planes;
      unsigned		  frames, width = 0, height = 0, n, bands, n_predicted = 0;
      unsigned		  n_intra = 0, has_b = 0;
      unsigned		 *order;		/* coding order */
      int		 *ctype;		/* frame types as coded */
      fb200_ctx_t	 *pctx = NULL;
      int16_t		**recon = NULL;		/* regenerated frames (references of P frames) */
      uint8_t		**delta = NULL;		/* delta flags of the states of P frames */
      const int16_t	**iplanes;
      int		  color = 0, rc;
      char		  err [512] = "";
      char		 *name;

      if (!inputname || !inputname [0] || strcmp (inputname [0], "-") == 0)
	 template = default_input;
      else
	 template = inputname;
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
	 default_options = fiasco_c_options_new ();
	 cop		 = fi_cast_c_options (default_options);
      }

      output = fi_bits_open (outputname);
      if (!output)
      {
	 fi_set_error ("Can't write outputfile `%s'.\n%s",
		       outputname ? outputname : "<stdout>", fi_system_error ());
	 if (default_options)
	    fiasco_c_options_delete (default_options);
	 return 0;
      }

      /* all frames readable, same size, same colour model (coder.c:204-240) */
      for (n = 0; (name = input_name (template, n)); n++)
      {
	 unsigned w, h;
	 int	  c;

	 fi_read_pnm_header (name, &w, &h, &c);
	 if (n)
	 {
	    if (w != width || h != height)
	       fi_error ("`%s': all images of a sequence have to be of the same size.", name);
	    if (c != color)
	       fi_error ("`%s': all images a sequence have to use the same color model.",
			 name);
	 }
	 else
	 {
	    width  = w;
	    height = h;
	    color  = c;
	 }
	 free (name);
      }
      frames = n;
      if (!frames)
	 fi_error ("No input frames.");
      bands = color ? 3 : 1;

      /* what this build does not do is refused, not approximated */
      order = fiasco_calloc (frames, sizeof (unsigned));
      ctype = fiasco_calloc (frames, sizeof (int));
      for (n = 0; n < frames; n++)
	 if (pattern_type (n, cop->pattern) < 0)
	    fi_error ("Frame type %c not valid. Choose one of I,B or P.",
		      cop->pattern [n % strlen (cop->pattern)]);
      coding_order (frames, cop->pattern, order, ctype);
      for (n = 0; n < frames; n++)
	 if (ctype [n])
	 {
	    if (ctype [n] == 2)
	       has_b = 1;
	    
	    if (color)
	       fi_error ("Predicted frames are available for grey sequences only: code colour "
			 "sequences with a frame pattern of I frames (--pattern=i).");
	    if (cop->half_pixel_prediction)
	       fi_error ("Half pixel motion compensation is not available in the B200 build.");
	    if (!cop->normal_domains || !cop->delta_domains
		|| cop->d_rpf_mantissa != cop->rpf_mantissa || cop->d_rpf_range != cop->rpf_range
		|| cop->d_dc_rpf_mantissa != cop->dc_rpf_mantissa
		|| cop->d_dc_rpf_range != cop->dc_rpf_range)
	       fi_error ("Predicted frames: only the default domain pool and quantisation "
			 "settings of the prediction errors are available.");
	    n_predicted++;
	 }
      if (cop->prediction)
	 fi_error ("Nondeterministic (DC) prediction is not available in the B200 build.");
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

      /* geometry and option clamping (coder.c:249-327) */
      memset (&p, 0, sizeof p);
      memset (&wi, 0, sizeof wi);
      {
	 unsigned lx = (unsigned) (log2 ((double) (width - 1)) + 1);
	 unsigned ly = (unsigned) (log2 ((double) (height - 1)) + 1);

	 wi.level = fi_max (lx, ly) * 2 - ((ly == lx + 1) ? 1 : 0);
      }
      p.width	     = (int) width;
      p.height	     = (int) height;
      p.bands	     = (int) bands;
      p.level	     = (int) wi.level;
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
      p.images_level	  = (int) fi_min (cop->images_level, (unsigned) p.lc_max_level - 1);
      wi.max_states	  = fi_max (fi_min (cop->max_states, FI_MAXSTATES), 1);
      p.max_states	  = (int) wi.max_states;
      p.max_elements	  = (int) fi_max (fi_min (cop->max_elements, FI_MAXEDGES), 1);
      wi.chroma_max_states = fi_max (1, cop->chroma_max_states);
      p.chroma_max_states = (int) wi.chroma_max_states;
      p.price		  = 128 * 64 / quality;		/* coder.c:164 */
      p.chroma_decrease	  = cop->chroma_decrease;
      wi.rpf	  = fi_make_rpf (cop->rpf_mantissa, (int) cop->rpf_range);
      wi.dc_rpf	  = fi_make_rpf (cop->dc_rpf_mantissa, (int) cop->dc_rpf_range);
      wi.d_rpf	  = fi_make_rpf (cop->d_rpf_mantissa, (int) cop->d_rpf_range);
      wi.d_dc_rpf = fi_make_rpf (cop->d_dc_rpf_mantissa, (int) cop->d_dc_rpf_range);
      p.rpf_mantissa	    = (int) wi.rpf.mantissa_bits;
      p.rpf_range	    = wi.rpf.range;
      p.dc_rpf_mantissa	    = (int) wi.dc_rpf.mantissa_bits;
      p.dc_rpf_range	    = wi.dc_rpf.range;
      p.second_domain_block = cop->second_domain_block;
      p.state_capacity	    = 0;
      wi.basis_name    = cop->basis_name;
      wi.title	       = cop->title;
      wi.comment       = cop->comment;
      wi.color	       = color;
      wi.width	       = width;
      wi.height	       = height;
      wi.frames	       = frames;
      wi.fps	       = cop->fps;
      wi.search_range  = cop->search_range;
      wi.half_pixel    = cop->half_pixel_prediction;
      wi.B_as_past_ref = cop->B_as_past_ref;
      wi.smoothing     = cop->smoothing;

      /* read every frame; intra frames are independent streams, so all of them go to
	 the device in one call (one thread block per frame) */
      images = fiasco_calloc (frames, sizeof (fi_image_t *));
      planes = fiasco_calloc ((size_t) frames * bands, sizeof (int16_t *));
      wfas   = fiasco_calloc (frames, sizeof (fb200_wfa_t));
      for (n = 0; n < frames; n++)
      {
	 unsigned b;

	 name	    = input_name (template, n);
	 images [n] = fi_read_image (name);
	 free (name);
	 for (b = 0; b < bands; b++)
	    planes [n * bands + b] = images [n]->pixels [b];
	 if (fb200_wfa_alloc (&wfas [n], FI_MAXSTATES))
	    fi_error ("Out of memory!");
      }
      /* the intra frames of the sequence: one launch */
      iplanes = fiasco_calloc ((size_t) frames * bands, sizeof (int16_t *));
      {
	 fb200_wfa_t *batch = fiasco_calloc (frames, sizeof (fb200_wfa_t));
	 unsigned     b, i;

	 for (n = 0; n < frames; n++)
	    if ((ctype [n] == 0))
	    {
	       for (b = 0; b < bands; b++)
		  iplanes [n_intra * bands + b] = planes [n * bands + b];
	       batch [n_intra++] = wfas [n];
	    }
	 rc = fb200_create (&ctx, &p, (int) n_intra, 0, err, sizeof err);
	 if (rc == FB200_OK)
	    rc = fb200_encode_tiles (ctx, (int) n_intra, iplanes, batch, NULL, 0, NULL, err,
				     sizeof err);
	 if (ctx)
	    fb200_destroy (ctx);
	 if (rc != FB200_OK)
	    fi_error ("%s", err [0] ? err : "GPU encoder failed");
	 for (n = 0, i = 0; n < frames; n++)
	    if ((ctype [n] == 0))
	       wfas [n] = batch [i++];
	 free (batch);
      }

      /*
       *  Predicted frames (video_coder, coder.c:490-680): a P frame needs the REGENERATED
       *  previous frame, so the frames of one group of pictures are a chain; the groups are
       *  independent.  Step k codes the k-th P frame of every group in one launch (one thread
       *  block per group), then the host closes the holes of the automata, derives the delta
       *  flags and regenerates the frames for step k + 1.
       */
      if (n_predicted && has_b)
      {
	 /*
	  *  Sequences with B frames: the frames in coding order, one launch per predicted frame,
	  *  with the reference bookkeeping of video_coder() (coder.c:571-627): a P frame is
	  *  predicted from the last regenerated frame; a B frame from a past and a future frame,
	  *  where the frame regenerated last becomes the future reference if it was coded ahead
	  *  of its display time, else (B frames serve as past references) the past one.
	  */
	 fb200_ctx_t	*bctx = NULL;
	 fb200_motion_t	 mo;
	 const int16_t	*past = NULL, *future = NULL, *cur = NULL;
	 int		 future_frame = 0, expected = 0;
	 unsigned	 k;
	 uint8_t	*seen = fiasco_calloc (frames + 1, 1);
	 jmp_buf	 saved;

	 recon = fiasco_calloc (frames, sizeof (int16_t *));
	 delta = fiasco_calloc (frames, sizeof (uint8_t *));
	 mo.p_min_level	 = (int) wi.p_min_level;
	 mo.p_max_level	 = (int) wi.p_max_level;
	 mo.search_range = (int) cop->search_range;
	 for (k = 0; k < frames; k++)
	 {
	    fiasco_frame_motion_t fm;
	    const int		  type = ctype [n = order [k]];

	    if (type == 1)
	    {
	       past   = cur;
	       future = NULL;
	    }
	    else if (type == 2)
	    {
	       if (future_frame)
		  future = cur;
	       else if (cop->B_as_past_ref)	/* else the last frame is dropped (coder.c:612-625) */
		  past = cur;
	    }
	    else
	       past = future = NULL;
	    seen [n]	 = 1;
	    future_frame = (int) n > expected;
	    while (expected < (int) frames && seen [expected])
	       expected++;
	    memset (&fm, 0, sizeof fm);
	    if (type)
	    {
	       fb200_ctx_t **cx = type == 2 ? &bctx : &pctx;

	       /* e.g. a B frame whose future reference is an I frame: the reference coder drops
		  the past frame there (coder.c:581-591) and then reads through the NULL pointer */
	       if (!past || (type == 2 && !future))
		  fi_error ("Frame %d (pattern `%s') has no reference frame to be predicted from.",
			    n, cop->pattern);
	       if (!*cx)
	       {
		  mo.frame_type = type;
		  rc = fb200_create_predicted (cx, &p, &mo, 1, 0, err, sizeof err);
		  if (rc != FB200_OK)
		     fi_error ("%s", err [0] ? err : "GPU encoder failed");
	       }
	       rc = fb200_encode_predicted (*cx, 1, &planes [n], &past, type == 2 ? &future : NULL,
					    &wfas [n], err, sizeof err);
	       if (rc != FB200_OK)
		  fi_error ("%s", err [0] ? err : "GPU encoder failed");
	       delta [n] = fiasco_calloc (FI_MAXSTATES, 1);
	       memcpy (saved, fi_env, sizeof saved);
	       rc = fiasco_finish_predicted_frame (&wfas [n], wfas [n].mv_type, wfas [n].mv_fx,
						   wfas [n].mv_fy, wfas [n].mv_bx, wfas [n].mv_by,
						   delta [n]);
	       memcpy (fi_env, saved, sizeof saved);
	       if (!rc)
		  fi_error ("%s", fiasco_get_error_message ());
	       fm.frame_type  = type;
	       fm.mv_type     = wfas [n].mv_type;
	       fm.mv_fx	      = wfas [n].mv_fx;
	       fm.mv_fy	      = wfas [n].mv_fy;
	       fm.mv_bx	      = wfas [n].mv_bx;
	       fm.mv_by	      = wfas [n].mv_by;
	       fm.delta_state = delta [n];
	    }
	    /* regenerate the frame: a reference of the frames to come (coder.c:642-651) */
	    recon [n] = fiasco_calloc ((size_t) width * height, sizeof (int16_t));
	    memcpy (saved, fi_env, sizeof saved);
	    rc = fiasco_regenerate_frame (&wfas [n], &fm, (int) width, (int) height, past, future,
					  recon [n]);
	    memcpy (fi_env, saved, sizeof saved);
	    if (!rc)
	       fi_error ("%s", fiasco_get_error_message ());
	    cur = recon [n];
	 }
	 if (pctx)
	    fb200_destroy (pctx);
	 if (bctx)
	    fb200_destroy (bctx);
	 free (seen);
      }
      else if (n_predicted)
      {
	 fb200_motion_t	 mo;
	 fb200_wfa_t	*batch	= fiasco_calloc (frames, sizeof (fb200_wfa_t));
	 const int16_t **bplane = fiasco_calloc (frames, sizeof (int16_t *));
	 const int16_t **bpast	= fiasco_calloc (frames, sizeof (int16_t *));
	 unsigned	*bframe = fiasco_calloc (frames, sizeof (unsigned));
	 unsigned	 k, groups = 0;
	 jmp_buf	 saved;

	 recon = fiasco_calloc (frames, sizeof (int16_t *));
	 delta = fiasco_calloc (frames, sizeof (uint8_t *));
	 for (n = 1; n < frames; n++)
	    if ((ctype [n] != 0) && (ctype [n - 1] == 0))
	       groups++;
	 mo.frame_type	 = 1;
	 mo.p_min_level	 = (int) wi.p_min_level;
	 mo.p_max_level	 = (int) wi.p_max_level;
	 mo.search_range = (int) cop->search_range;
	 rc = fb200_create_predicted (&pctx, &p, &mo, (int) groups, 0, err, sizeof err);
	 if (rc != FB200_OK)
	    fi_error ("%s", err [0] ? err : "GPU encoder failed");
	 for (k = 1; ; k++)
	 {
	    unsigned cnt = 0, i;

	    for (n = k; n < frames; n++)
	    {
	       unsigned j;

	       if ((ctype [n] == 0))
		  continue;
	       for (j = 1; j < k && (ctype [n - j] != 0); j++)
		  ;
	       if (j != k || (ctype [n - k] != 0))
		  continue;			/* not the k-th frame of its group */
	       if (!recon [n - 1])
	       {
		  /* the reference frame: regenerate the frame before (coder.c:642-651) */
		  fiasco_frame_motion_t fm;

		  memset (&fm, 0, sizeof fm);
		  recon [n - 1] = fiasco_calloc ((size_t) width * height, sizeof (int16_t));
		  if (k > 1)
		  {
		     fm.frame_type  = 1;
		     fm.mv_type	    = wfas [n - 1].mv_type;
		     fm.mv_fx	    = wfas [n - 1].mv_fx;
		     fm.mv_fy	    = wfas [n - 1].mv_fy;
		     fm.delta_state = delta [n - 1];
		  }
		  memcpy (saved, fi_env, sizeof saved);
		  rc = fiasco_regenerate_frame (&wfas [n - 1], &fm, (int) width, (int) height,
						k > 1 ? recon [n - 2] : NULL, NULL, recon [n - 1]);
		  memcpy (fi_env, saved, sizeof saved);
		  if (!rc)
		     fi_error ("%s", fiasco_get_error_message ());
	       }
	       bplane [cnt] = planes [n];
	       bpast [cnt]  = recon [n - 1];
	       bframe [cnt] = n;
	       batch [cnt]  = wfas [n];
	       cnt++;
	    }
	    if (!cnt)
	       break;
	    rc = fb200_encode_predicted (pctx, (int) cnt, bplane, bpast, NULL, batch, err, sizeof err);
	    if (rc != FB200_OK)
	    {
	       fb200_destroy (pctx);
	       fi_error ("%s", err [0] ? err : "GPU encoder failed");
	    }
	    for (i = 0; i < cnt; i++)
	    {
	       n	 = bframe [i];
	       wfas [n]	 = batch [i];
	       delta [n] = fiasco_calloc (FI_MAXSTATES, 1);
	       memcpy (saved, fi_env, sizeof saved);
	       rc = fiasco_finish_predicted_frame (&wfas [n], wfas [n].mv_type, wfas [n].mv_fx,
						   wfas [n].mv_fy, NULL, NULL, delta [n]);
	       memcpy (fi_env, saved, sizeof saved);
	       if (!rc)
		  fi_error ("%s", fiasco_get_error_message ());
	    }
	 }
	 fb200_destroy (pctx);
	 free (batch);
	 free (bplane);
	 free (bpast);
	 free (bframe);
      }
      free (iplanes);

      for (unsigned coded = 0; coded < frames; coded++)
      {
	 fi_wfa_t w;

	 n = order [coded];		/* the stream holds the frames in coding order */
	 memset (&w, 0, sizeof w);	/* intra frame: no motion data */
	 if ((ctype [n] != 0))
	 {
	    w.frame_type  = ctype [n];
	    w.mv_bx	  = (const int8_t (*)[2]) wfas [n].mv_bx;
	    w.mv_by	  = (const int8_t (*)[2]) wfas [n].mv_by;
	    w.x		  = (const uint16_t (*)[2]) wfas [n].x;
	    w.y		  = (const uint16_t (*)[2]) wfas [n].y;
	    w.mv_type	  = (const int8_t (*)[2]) wfas [n].mv_type;
	    w.mv_fx	  = (const int8_t (*)[2]) wfas [n].mv_fx;
	    w.mv_fy	  = (const int8_t (*)[2]) wfas [n].mv_fy;
	    w.delta_state = delta [n];
	 }
	 w.info		  = &wi;
	 w.states	  = wfas [n].states;
	 w.basis_states	  = wfas [n].basis_states;
	 w.root_state	  = wfas [n].root_state;
	 w.level_of_state = wfas [n].level_of_state;
	 w.domain_type	  = wfas [n].domain_type;
	 w.tree		  = (const int16_t (*)[2]) wfas [n].tree;
	 w.into		  = (const int16_t (*)[2][6]) wfas [n].into;
	 w.weight	  = (const float (*)[2][6]) wfas [n].weight;
	 w.y_state	  = (const int16_t (*)[2]) wfas [n].y_state;
	 w.y_column	  = (const uint8_t (*)[2]) wfas [n].y_column;
	 fi_debug_message ("WFA contains %d states (%d basis states).", w.states,
			   w.basis_states);
	 fi_debug_message ("Total costs : %.2f", (double) wfas [n].costs [0]);
	 fi_write_next_wfa (&w, n, coded == 0, cop->normal_domains, cop->delta_domains, output);
	 fb200_wfa_free (&wfas [n]);
	 fi_free_image (images [n]);
	 if (recon)
	    free (recon [n]);
	 if (delta)
	    free (delta [n]);
      }
      free (recon);
      free (delta);
      free (order);
      free (ctype);
      if (cop->progress_meter != FIASCO_PROGRESS_NONE)
	 fi_message ("");
      fi_bits_close (output);
      free (images);
      free (planes);
      free (wfas);
      if (default_options)
	 fiasco_c_options_delete (default_options);
      return 1;
   }
   fi_catch
   {
      return 0;
   }
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
   fi_try
   {
      fi_wfainfo_t wi;
      fi_bits_t	  *out;
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
	    if (motion [n].frame_type < 1 || motion [n].frame_type > 2 || info->color)
	       fi_error ("frame %d: only grey P and B frames can be written", n);
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
	 fi_write_next_wfa (&w, motion ? (unsigned) motion [n].frame_number : (unsigned) n, n == 0, 1, 1, out);
      }
      fi_bits_close (out);
      return 1;
   }
   fi_catch
   {
      return 0;
   }
}
