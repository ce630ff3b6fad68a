/*
 *  regenerate.c -- the image an automaton describes, in the coder's pixel format.
 *
 *  A coder of sequences regenerates every frame it has coded: the result is the reference
 *  of the next predicted frame (reference: codec/coder.c:642-651 -- decode_image,
 *  codec/decoder.c:412-535 with alloc_state_images :878 and compute_state_images :1107;
 *  restore_mc, codec/motion.c:37-230 with extract_mc_block :232).  Host side of the motion
 *  path; integer arithmetic throughout, so it is exact by
 *  construction: the reference adds two pixels per 32-bit int with one guard bit each,
 *  which per pixel is a wrapping 16-bit addition of even numbers.
 */
#include <stdlib.h>
#include <string.h>

#include "fiasco_host.h"
#include "fi_internal.h"

#define W_OF(l) (1u << ((l) >> 1))
#define H_OF(l) (1u << (((l) + 1) >> 1))

typedef struct regen
{
   const fb200_wfa_t *w;
   int16_t	    **pix;	/* [states * (max_level + 1)] image of a state at a level */
   unsigned	      levels;
} regen_t;

static const int16_t *
state_image (regen_t *r, unsigned state, unsigned level)
{
   const fb200_wfa_t *w	   = r->w;
   int16_t	    **slot = &r->pix [(size_t) state * r->levels + level];

   if (*slot)
      return *slot;
   int16_t *img = fiasco_calloc ((size_t) 1 << level, sizeof (int16_t));

   *slot = img;
   if (level == 0)			/* decoder.c:1128-1130 */
   {
      img [0] = (int16_t) ((int) (w->final_distribution [state] * 8 + .5) * 2);
      return img;
   }
   const unsigned width = W_OF (level - 1), height = H_OF (level - 1), stride = W_OF (level);

   for (unsigned label = 0; label < 2; label++)
   {
      /* odd levels split into an upper and a lower half, even ones into left and right */
      int16_t  *range = (level & 1) ? img + label * height * stride : img + label * width;
      const int child = w->tree [2 * state + label];

      if (child >= 0)
      {
	 const int16_t *src = state_image (r, (unsigned) child, level - 1);

	 for (unsigned y = 0; y < height; y++)
	    memcpy (range + y * stride, src + y * width, width * sizeof (int16_t));
      }
      for (unsigned e = 0; w->into [(2 * state + label) * 6 + e] >= 0; e++)
      {
	 const int   domain = w->into [(2 * state + label) * 6 + e];
	 const float weight = w->weight [(2 * state + label) * 6 + e];

	 if (domain != 0)
	 {
	    const int16_t *src = state_image (r, (unsigned) domain, level - 1);
	    const int	   iw  = (int16_t) (weight * 512 + 0.5);	/* wfalib.c:273 */

	    for (unsigned y = 0; y < height; y++)
	       for (unsigned x = 0; x < width; x++)
		  range [y * stride + x]
		     = (int16_t) (range [y * stride + x] + (((iw * (int) src [y * width + x]) >> 10) << 1));
	 }
	 else			/* the constant state: one value for the whole range */
	 {
	    const int c = (int) (weight * w->final_distribution [0] * 8 + .5) * 2;

	    for (unsigned y = 0; y < height; y++)
	       for (unsigned x = 0; x < width; x++)
		  range [y * stride + x] = (int16_t) (range [y * stride + x] + c);
	 }
      }
   }
   return img;
}

/* grey: one plane; colour (4:4:4, what the coder regenerates: coder.c:647): three planes Y, Cb, Cr
   of width * height shorts behind each other in out / past / future */
static int
regenerate (const fb200_wfa_t *w, const fiasco_frame_motion_t *motion, int width, int height,
	    int colour, const int16_t *past, const int16_t *future, int16_t *out)
{
   fi_try
   {
      regen_t	   r;
      unsigned	   max_level = 0, aw = 0, ah = 0, state;
      const size_t npix	 = (size_t) width * (size_t) height;
      unsigned	   root [2] = {0, 0};	/* colour: the roots of the Y and the Cb band (decoder.c:436-444) */

      if (!w || !out || width < 1 || height < 1 || w->status != FB200_OK)
      {
	 fi_set_error ("fiasco_regenerate_frame: bad arguments");
	 return 0;
      }
      if (motion && motion->frame_type != 0
	  && (motion->frame_type > 2 || !past || (motion->frame_type == 2 && !future)))
      {
	 fi_set_error ("fiasco_regenerate_frame: a predicted frame needs its reference frame(s)");
	 return 0;
      }
      /* highest level of a linear combination; the size the bintree covers (decoder.c:449-461,
	 843-875) */
      for (state = w->basis_states; state < w->states; state++)
	 if (w->into [(2 * state) * 6] >= 0 || w->into [(2 * state + 1) * 6] >= 0)
	 {
	    const unsigned l = w->level_of_state [state];

	    if (l > max_level)
	       max_level = l;
	    if (w->x [2 * state] + W_OF (l) > aw)
	       aw = w->x [2 * state] + W_OF (l);
	    if (w->y [2 * state] + H_OF (l) > ah)
	       ah = w->y [2 * state] + H_OF (l);
	 }
      aw += aw & 1;
      ah += ah & 1;
      if (aw < (unsigned) width)
	 aw = (unsigned) width;
      if (ah < (unsigned) height)
	 ah = (unsigned) height;
      if (colour)
      {
	 const int top0 = w->tree [2 * w->root_state], top1 = w->tree [2 * w->root_state + 1];

	 if (top0 < 0 || top1 < 0 || w->tree [2 * top0] < 0 || w->tree [2 * top0 + 1] < 0)
	    fi_error ("fiasco_regenerate_frame: not the automaton of a colour frame");
	 root [0] = (unsigned) w->tree [2 * top0];
	 root [1] = (unsigned) w->tree [2 * top0 + 1];
      }
      r.w      = w;
      r.levels = max_level + 1;
      r.pix    = fiasco_calloc ((size_t) w->states * r.levels, sizeof (int16_t *));

      const unsigned bands = colour ? 3 : 1;
      int16_t	    *frame = fiasco_calloc ((size_t) aw * ah * bands, sizeof (int16_t));

      /* every state of level max_level is one block of the frame (decoder.c:913-937); the states
	 of a colour frame's bands follow each other: Y up to its root, then Cb, then Cr */
      for (state = w->basis_states; state < w->states; state++)
	 if (w->level_of_state [state] == max_level)
	 {
	    const unsigned b  = !colour || state <= root [0] ? 0 : state > root [1] ? 2 : 1;
	    const unsigned bw = W_OF (max_level), bh = H_OF (max_level);
	    const unsigned x0 = w->x [2 * state], y0 = w->y [2 * state];
	    const unsigned cw = x0 >= aw ? 0 : (aw - x0 < bw ? aw - x0 : bw);
	    const int16_t *img = state_image (&r, state, max_level);

	    for (unsigned y = 0; y < bh && y0 + y < ah; y++)
	       memcpy (frame + (size_t) b * aw * ah + (size_t) (y0 + y) * aw + x0, img + y * bw,
		       cw * sizeof (int16_t));
	 }
      for (unsigned b = 0; b < bands; b++)
	 for (int y = 0; y < height; y++)
	    memcpy (out + b * npix + (size_t) y * width, frame + (size_t) b * aw * ah + (size_t) y * aw,
		    (size_t) width * sizeof (int16_t));
      free (frame);
      for (size_t i = 0; i < (size_t) w->states * r.levels; i++)
	 free (r.pix [i]);
      free (r.pix);

      /* restore_mc (motion.c:37-190) with full-pixel vectors: add the displaced block of the
	 previous frame (forward), of the future frame (backward), or their mean (interpolated,
	 arithmetic shift) */
      if (motion && motion->frame_type != 0)
	 for (state = w->basis_states; state <= (colour ? root [0] : w->root_state); state++)
	    for (unsigned label = 0; label < 2; label++)
	    {
	       const int type = motion->mv_type [2 * state + label];

	       if (type == 0)
		  continue;
	       if (type > 1 && (!future || !motion->mv_bx || !motion->mv_by))
		  fi_error ("backward motion compensation without a future frame");
	       const unsigned level = (unsigned) w->level_of_state [state] - 1;
	       const unsigned bw = W_OF (level), bh = H_OF (level);
	       const int      x0 = w->x [2 * state + label], y0 = w->y [2 * state + label];
	       const int      fx = motion->mv_fx [2 * state + label], fy = motion->mv_fy [2 * state + label];
	       const int      bx = type > 1 ? motion->mv_bx [2 * state + label] : 0;
	       const int      by = type > 1 ? motion->mv_by [2 * state + label] : 0;

	       /* (a colour frame: the luminance tree's vectors move all three bands, motion.c:59-62) */
	       for (unsigned band = 0; band < bands; band++)
		  for (unsigned y = 0; y < bh; y++)
		     for (unsigned x = 0; x < bw; x++)
		     {
			int16_t	 *o = out + band * npix + (size_t) (y0 + (int) y) * width + x0 + (int) x;
			const int f = type != 2 ? past [band * npix + (size_t) (y0 + fy + (int) y) * width + x0 + fx + (int) x] : 0;
			const int b = type != 1 ? future [band * npix + (size_t) (y0 + by + (int) y) * width + x0 + bx + (int) x] : 0;

			*o = (int16_t) (*o + (type == 1 ? f : type == 2 ? b : (f + b) >> 1));
		     }
	    }
      /* the chroma bands of a predicted colour frame are clipped to 8 bits (motion.c:192-224) */
      if (colour && motion && motion->frame_type != 0)
	 for (size_t n = npix; n < 3 * npix; n++)
	 {
	    int v = out [n] >> 4;

	    v	    = v < -128 ? -128 : v > 127 ? 127 : v;
	    out [n] = (int16_t) (v * 16);
	 }
      return 1;
   }
   fi_catch
   {
      return 0;
   }
}

int
fiasco_regenerate_frame (const fb200_wfa_t *w, const fiasco_frame_motion_t *motion,
			 int width, int height, const int16_t *past, const int16_t *future,
			 int16_t *out)
{
   return regenerate (w, motion, width, height, 0, past, future, out);
}

int
fiasco_regenerate_colour_frame (const fb200_wfa_t *w, const fiasco_frame_motion_t *motion,
				int width, int height, const int16_t *past, const int16_t *future,
				int16_t *out)
{
   return regenerate (w, motion, width, height, 1, past, future, out);
}

/*
 *  Finish the automaton of a predicted frame as the device leaves it:
 *  the states the losing split alternatives left behind are holes (level_of_state == 255);
 *  close them by a monotone renumbering, then derive the delta flags from the structure
 *  (reference: locate_delta_images, codec/wfalib.c:699-730, called at codec/coder.c:876).
 *  Works in place on the automaton's arrays and the motion arrays; returns the new number
 *  of states.
 */
int
fiasco_finish_predicted_frame (fb200_wfa_t *w, int8_t *mv_type, int8_t *mv_fx, int8_t *mv_fy,
			       int8_t *mv_bx, int8_t *mv_by, uint8_t *delta_state)
{
   fi_try
   {
      if (!w || !mv_type || !mv_fx || !mv_fy || !delta_state || w->status != FB200_OK)
      {
	 fi_set_error ("fiasco_finish_predicted_frame: bad arguments");
	 return 0;
      }
      int16_t *map = fiasco_calloc (w->states + 1, sizeof (int16_t));
      unsigned n   = 0, s;

      for (s = 0; s < w->states; s++)
	 map [s] = (s >= w->basis_states && w->level_of_state [s] == 255) ? (int16_t) -1 : (int16_t) n++;
      for (s = 0; s < w->states; s++)
      {
	 const int t = map [s];

	 if (t < 0 || (unsigned) t == s)
	    continue;
	 w->final_distribution [t] = w->final_distribution [s];
	 w->level_of_state [t]	   = w->level_of_state [s];
	 w->domain_type [t]	   = w->domain_type [s];
	 for (unsigned label = 0; label < 2; label++)
	 {
	    const unsigned a = 2 * (unsigned) t + label, b = 2 * s + label;

	    w->tree [a]	    = w->tree [b];
	    w->x [a]	    = w->x [b];
	    w->y [a]	    = w->y [b];
	    w->y_state [a]  = w->y_state [b];
	    w->y_column [a] = w->y_column [b];
	    mv_type [a]	    = mv_type [b];
	    mv_fx [a]	    = mv_fx [b];
	    mv_fy [a]	    = mv_fy [b];
	    if (mv_bx && mv_by)
	    {
	       mv_bx [a] = mv_bx [b];
	       mv_by [a] = mv_by [b];
	    }
	    memcpy (w->into + a * 6, w->into + b * 6, 6 * sizeof (int16_t));
	    memcpy (w->weight + a * 6, w->weight + b * 6, 6 * sizeof (float));
	 }
      }
      for (s = 0; s < n; s++)
	 for (unsigned label = 0; label < 2; label++)
	 {
	    const unsigned a = 2 * s + label;

	    if (w->tree [a] >= 0)
	       w->tree [a] = map [w->tree [a]];
	    if (w->y_state [a] >= 0)		/* (chroma bands of a colour frame) */
	       w->y_state [a] = map [w->y_state [a]];
	    for (unsigned e = 0; w->into [a * 6 + e] >= 0; e++)
	       w->into [a * 6 + e] = map [w->into [a * 6 + e]];
	 }
      w->root_state = (unsigned) map [w->root_state];
      w->states	    = n;
      free (map);

      /* locate_delta_images: top down, a child is a delta state if its range is motion
	 compensated, or has edges beside the child, or its parent is a delta state */
      for (s = w->basis_states; s < n; s++)
	 delta_state [s] = 0;
      for (s = w->root_state + 1; s-- > w->basis_states; )
	 for (unsigned label = 0; label < 2; label++)
	 {
	    const unsigned a = 2 * s + label;

	    if (w->tree [a] >= 0 && (mv_type [a] != 0 || w->into [a * 6] >= 0 || delta_state [s]))
	       delta_state [w->tree [a]] = 1;
	 }
      return (int) n;
   }
   fi_catch
   {
      return 0;
   }
}
