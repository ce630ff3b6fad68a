/*
 *  options.c -- fiasco_c_options_*: the coder's option object.
 *  Defaults, argument checks and error texts follow the reference (codec/options.c:29-706):
 *  library defaults are levels [4,12] with 5 edges (the command line front end overrides
 *  them with [6,10] / 3 edges for -z 0, bin/cwfa.c:326-345).
 */
#include <stdlib.h>
#include <string.h>

#include "fi_internal.h"

fiasco_c_options_t *
fiasco_c_options_new (void)
{
   c_options_t	      *o   = fiasco_calloc (1, sizeof (c_options_t));
   fiasco_c_options_t *pub = fiasco_calloc (1, sizeof (fiasco_c_options_t));

   pub->private		   = o;
   pub->delete		   = fiasco_c_options_delete;
   pub->set_tiling	   = fiasco_c_options_set_tiling;
   pub->set_frame_pattern  = fiasco_c_options_set_frame_pattern;
   pub->set_basisfile	   = fiasco_c_options_set_basisfile;
   pub->set_chroma_quality = fiasco_c_options_set_chroma_quality;
   pub->set_optimizations  = fiasco_c_options_set_optimizations;
   /* the reference leaves this one method pointer NULL (options.c:47-59); callers use
      the free function.  Filling it in is harmless and is what the header promises. */
   pub->set_prediction	   = fiasco_c_options_set_prediction;
   pub->set_video_param	   = fiasco_c_options_set_video_param;
   pub->set_quantization   = fiasco_c_options_set_quantization;
   pub->set_progress_meter = fiasco_c_options_set_progress_meter;
   pub->set_smoothing	   = fiasco_c_options_set_smoothing;
   pub->set_title	   = fiasco_c_options_set_title;
   pub->set_comment	   = fiasco_c_options_set_comment;

   strcpy (o->id, "COFIASCO");
   o->basis_name	    = strdup ("small.fco");
   o->lc_min_level	    = 4;
   o->lc_max_level	    = 12;
   o->p_min_level	    = 8;
   o->p_max_level	    = 10;
   o->images_level	    = 5;
   o->max_states	    = FI_MAXSTATES;
   o->chroma_max_states	    = 40;
   o->max_elements	    = FI_MAXEDGES;
   o->tiling_exponent	    = 4;
   o->tiling_method	    = FIASCO_TILING_VARIANCE_DSC;
   o->id_domain_pool	    = strdup ("rle");
   o->id_d_domain_pool	    = strdup ("rle");
   o->id_rpf_model	    = strdup ("adaptive");
   o->id_d_rpf_model	    = strdup ("adaptive");
   o->rpf_mantissa	    = 3;
   o->rpf_range		    = FIASCO_RPF_RANGE_1_50;
   o->dc_rpf_mantissa	    = 5;
   o->dc_rpf_range	    = FIASCO_RPF_RANGE_1_00;
   o->d_rpf_mantissa	    = 3;
   o->d_rpf_range	    = FIASCO_RPF_RANGE_1_50;
   o->d_dc_rpf_mantissa	    = 5;
   o->d_dc_rpf_range	    = FIASCO_RPF_RANGE_1_00;
   o->chroma_decrease	    = 2.0;
   o->prediction	    = 0;
   o->delta_domains	    = 1;
   o->normal_domains	    = 1;
   o->search_range	    = 16;
   o->fps		    = 25;
   o->pattern		    = strdup ("IPPPPPPPPP");
   o->reference_filename    = NULL;
   o->half_pixel_prediction = 0;
   o->cross_B_search	    = 1;
   o->B_as_past_ref	    = 1;
   o->progress_meter	    = FIASCO_PROGRESS_NONE;
   o->smoothing		    = 70;
   o->comment		    = strdup ("");
   o->title		    = strdup ("");
   return pub;
}

c_options_t *
fi_cast_c_options (const fiasco_c_options_t *options)
{
   c_options_t *o = options ? (c_options_t *) options->private : NULL;

   if (o)
   {
      if (strcmp (o->id, "COFIASCO") != 0)
      {
	 fi_set_error ("Parameter `options' doesn't match required type.");
	 return NULL;
      }
   }
   else
      fi_set_error ("Parameter `%s' not defined (NULL).", "options");
   return o;
}

void
fiasco_c_options_delete (fiasco_c_options_t *options)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return;
   free (o->basis_name);
   free (o->id_domain_pool);
   free (o->id_d_domain_pool);
   free (o->id_rpf_model);
   free (o->id_d_rpf_model);
   free (o->pattern);
   free (o->comment);
   free (o->title);
   free (o);
   free (options);
}

int
fiasco_c_options_set_tiling (fiasco_c_options_t *options, fiasco_tiling_e method,
			     unsigned exponent)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   switch (method)
   {
      case FIASCO_TILING_SPIRAL_ASC:
      case FIASCO_TILING_SPIRAL_DSC:
      case FIASCO_TILING_VARIANCE_ASC:
      case FIASCO_TILING_VARIANCE_DSC:
	 o->tiling_method = method;
	 break;
      default:
	 fi_set_error ("Invalid tiling method `%d' specified "
		       "(valid methods are 0, 1, 2, or 3).", method);
	 return 0;
   }
   o->tiling_exponent = exponent;
   return 1;
}

int
fiasco_c_options_set_frame_pattern (fiasco_c_options_t *options, const char *pattern)
{
   c_options_t *o = fi_cast_c_options (options);
   const char	*s;

   if (!o)
      return 0;
   if (!pattern)
   {
      fi_set_error ("Parameter `%s' not defined (NULL).", "pattern");
      return 0;
   }
   if (strlen (pattern) < 1)
   {
      fi_set_error ("Frame type pattern doesn't contain any character.");
      return 0;
   }
   for (s = pattern; *s; s++)
      if (!strchr ("iIbBpP", *s))
      {
	 fi_set_error ("Frame type pattern contains invalid character `%c' "
		       "(choose I, B or P).", *s);
	 return 0;
      }
   free (o->pattern);
   o->pattern = strdup (pattern);
   return 1;
}

int
fiasco_c_options_set_basisfile (fiasco_c_options_t *options, const char *filename)
{
   c_options_t *o = fi_cast_c_options (options);
   FILE	       *file;

   if (!o)
      return 0;
   if (!filename)
   {
      fi_set_error ("Parameter `%s' not defined (NULL).", "filename");
      return 0;
   }
   file = open_file (filename, "FIASCO_DATA", READ_ACCESS);
   if (!file)
   {
      fi_set_error ("Can't read basis file `%s'.\n%s.", filename, fi_system_error ());
      return 0;
   }
   fclose (file);
   free (o->basis_name);
   o->basis_name = strdup (filename);
   return 1;
}

int
fiasco_c_options_set_chroma_quality (fiasco_c_options_t *options, float quality_factor,
				     unsigned dictionary_size)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   if (!dictionary_size)
   {
      fi_set_error ("Size of chroma compression dictionary has to be a positive number.");
      return 0;
   }
   if (quality_factor <= 0)
   {
      fi_set_error ("Quality of chroma channel compression has to be positive value.");
      return 0;
   }
   o->chroma_decrease	= quality_factor;
   o->chroma_max_states = dictionary_size;
   return 1;
}

int
fiasco_c_options_set_optimizations (fiasco_c_options_t *options, unsigned min_block_level,
				    unsigned max_block_level, unsigned max_elements,
				    unsigned dictionary_size, unsigned optimization_level)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   if (!dictionary_size)
   {
      fi_set_error ("Size of dictionary has to be a positive number.");
      return 0;
   }
   if (!max_elements)
   {
      fi_set_error ("At least one dictionary element has to be used in an approximation.");
      return 0;
   }
   if (max_block_level < 4)
   {
      fi_set_error ("Maximum image block size has to be at least level 4.");
      return 0;
   }
   if (min_block_level < 4)
   {
      fi_set_error ("Minimum image block size has to be at least level 4.");
      return 0;
   }
   if (max_block_level < min_block_level)
   {
      fi_set_error ("Maximum block size has to be larger or equal minimum block size.");
      return 0;
   }
   o->lc_min_level	  = min_block_level;
   o->lc_max_level	  = max_block_level;
   o->max_states	  = dictionary_size;
   o->max_elements	  = max_elements;
   o->second_domain_block = optimization_level > 0;
   o->check_for_overflow  = optimization_level > 1;
   o->check_for_underflow = optimization_level > 1;
   o->full_search	  = optimization_level > 1;
   return 1;
}

int
fiasco_c_options_set_prediction (fiasco_c_options_t *options, int intra_prediction,
				 unsigned min_block_level, unsigned max_block_level)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   if (max_block_level < 6)
   {
      fi_set_error ("Maximum prediction block size has to be at least level 6");
      return 0;
   }
   if (min_block_level < 6)
   {
      fi_set_error ("Minimum prediction block size has to be at least level 6");
      return 0;
   }
   if (max_block_level < min_block_level)
   {
      fi_set_error ("Maximum prediction block size has to be larger or "
		    "equal minimum block size.");
      return 0;
   }
   o->p_min_level = min_block_level;
   o->p_max_level = max_block_level;
   o->prediction  = intra_prediction;
   return 1;
}

int
fiasco_c_options_set_video_param (fiasco_c_options_t *options, unsigned frames_per_second,
				  int half_pixel_prediction, int cross_B_search,
				  int B_as_past_ref)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   o->fps		    = frames_per_second;
   o->half_pixel_prediction = half_pixel_prediction;
   o->cross_B_search	    = cross_B_search;
   o->B_as_past_ref	    = B_as_past_ref;
   return 1;
}

static int
valid_range (fiasco_rpf_range_e r)
{
   return r == FIASCO_RPF_RANGE_0_75 || r == FIASCO_RPF_RANGE_1_00
	  || r == FIASCO_RPF_RANGE_1_50 || r == FIASCO_RPF_RANGE_2_00;
}

int
fiasco_c_options_set_quantization (fiasco_c_options_t *options, unsigned mantissa,
				   fiasco_rpf_range_e range, unsigned dc_mantissa,
				   fiasco_rpf_range_e dc_range)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   if (mantissa < 2 || mantissa > 8 || dc_mantissa < 2 || dc_mantissa > 8)
   {
      fi_set_error ("Number of RPF mantissa bits `%d', `%d' have to be in "
		    "the interval [2,8].", mantissa, dc_mantissa);
      return 0;
   }
   if (!valid_range (range) || !valid_range (dc_range))
   {
      fi_set_error ("Invalid RPF ranges `%d', `%d' specified.", range, dc_range);
      return 0;
   }
   o->rpf_range	      = range;
   o->dc_rpf_range    = dc_range;
   o->rpf_mantissa    = mantissa;
   o->dc_rpf_mantissa = dc_mantissa;
   return 1;
}

int
fiasco_c_options_set_progress_meter (fiasco_c_options_t *options, fiasco_progress_e type)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   switch (type)
   {
      case FIASCO_PROGRESS_BAR:
      case FIASCO_PROGRESS_PERCENT:
      case FIASCO_PROGRESS_NONE:
	 o->progress_meter = type;
	 break;
      default:
	 fi_set_error ("Invalid progress meter `%d' specified "
		       "(valid values are 0, 1, or 2).", type);
	 return 0;
   }
   return 1;
}

int
fiasco_c_options_set_smoothing (fiasco_c_options_t *options, int smoothing)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   if (smoothing < -1 || smoothing > 100)
   {
      fi_set_error ("Smoothing percentage must be in the range [-1, 100].");
      return 0;
   }
   o->smoothing = smoothing;
   return 1;
}

int
fiasco_c_options_set_comment (fiasco_c_options_t *options, const char *comment)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   if (!comment)
   {
      fi_set_error ("Parameter `%s' not defined (NULL).", "title");
      return 0;
   }
   free (o->comment);
   o->comment = strdup (comment);
   return 1;
}

int
fiasco_c_options_set_title (fiasco_c_options_t *options, const char *title)
{
   c_options_t *o = fi_cast_c_options (options);

   if (!o)
      return 0;
   if (!title)
   {
      fi_set_error ("Parameter `%s' not defined (NULL).", "title");
      return 0;
   }
   free (o->title);
   o->title = strdup (title);
   return 1;
}
