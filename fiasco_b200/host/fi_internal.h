/*
 *  fi_internal.h -- internals of the host side of libfiasco (B200 build): error handling,
 *  options, PNM input, bit stream output.  Plain C; the hot path is reached only through
 *  the C ABI of include/fiasco_b200.h.
 */
#ifndef FI_INTERNAL_H
#define FI_INTERNAL_H

#include <setjmp.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "fiasco.h"
#include "fiasco_b200.h"

#define FI_MAXEDGES  FB200_MAXEDGES
#define FI_MAXSTATES FB200_MAXSTATES
#define FI_MAXLABELS FB200_MAXLABELS
#define FI_MAXLEVEL  FB200_MAXLEVEL
#define FI_MAXSTRLEN 1024
#define FI_NO_EDGE   (-1)
#define FI_RANGE     (-1)
#define FI_USE_DOMAIN 2

/* ---- errors: same try/catch discipline as the reference (lib/error.h:43-45) ---- */
extern __thread jmp_buf fi_env;
#define fi_try	 if (setjmp (fi_env) == 0)
#define fi_catch else
void fi_set_error (const char *format, ...);
void fi_error (const char *format, ...);		/* set text, longjmp */
void fi_rethrow (void);				/* longjmp, the stored text stays */
void fi_file_error (const char *filename);
void fi_warning (const char *format, ...);
void fi_message (const char *format, ...);
void fi_info (const char *format, ...);
void fi_debug_message (const char *format, ...);
const char *fi_system_error (void);

/* ---- memory / files that the reference CLI links directly (lib/misc.c:51, bit-io.c:48) */
typedef enum {READ_ACCESS, WRITE_ACCESS} openmode_e;
void *fiasco_calloc (size_t n, size_t size);
void  fiasco_free (void *ptr);
FILE *open_file (const char *filename, const char *env_var, openmode_e mode);
/* fn (i, ctx) for i in 0..n-1 on up to FIASCO_HOST_THREADS host threads (default: the cores of the
   process, at most 16); returns 0, or 1 if a call failed through fi_error (the message is kept) */
int   fi_parallel_for (size_t n, void (*fn) (size_t i, void *ctx), void *ctx);

/* ---- options (codec/options.h:20-65) ---- */
typedef struct c_options
{
   char		      id [9];
   char		     *basis_name;
   unsigned	      lc_min_level, lc_max_level;
   unsigned	      p_min_level, p_max_level;
   unsigned	      images_level;
   unsigned	      max_states, chroma_max_states, max_elements;
   unsigned	      tiling_exponent;
   fiasco_tiling_e    tiling_method;
   char		     *id_domain_pool, *id_d_domain_pool, *id_rpf_model, *id_d_rpf_model;
   unsigned	      rpf_mantissa;
   fiasco_rpf_range_e rpf_range;
   unsigned	      dc_rpf_mantissa;
   fiasco_rpf_range_e dc_rpf_range;
   unsigned	      d_rpf_mantissa;
   fiasco_rpf_range_e d_rpf_range;
   unsigned	      d_dc_rpf_mantissa;
   fiasco_rpf_range_e d_dc_rpf_range;
   float	      chroma_decrease;
   int		      prediction, delta_domains, normal_domains;
   unsigned	      search_range, fps;
   char		     *pattern;
   char		     *reference_filename;
   int		      half_pixel_prediction, cross_B_search, B_as_past_ref;
   int		      check_for_underflow, check_for_overflow, second_domain_block,
		      full_search;
   fiasco_progress_e  progress_meter;
   char		     *title, *comment;
   unsigned	      smoothing;
} c_options_t;

c_options_t *fi_cast_c_options (const fiasco_c_options_t *options);

/* ---- PNM input (lib/image.c:262-388) ---- */
typedef struct fi_image
{
   unsigned width, height;
   int	    color;
   int16_t *pixels [3];		/* 12.4 fixed point: grey, or Y Cb Cr (4:4:4) */
} fi_image_t;

void	    fi_read_pnm_header (const char *name, unsigned *width, unsigned *height,
				int *color);
fi_image_t *fi_read_image (const char *name);
void	    fi_free_image (fi_image_t *image);

/* ---- reduced precision format (lib/rpf.c) -- host copy for the stream writer ---- */
typedef struct fi_rpf
{
   unsigned mantissa_bits;
   float    range;
   int	    range_e;
} fi_rpf_t;
fi_rpf_t fi_make_rpf (unsigned mantissa, int range_e);
int	 fi_rtob (float f, const fi_rpf_t *rpf);

/* ---- bit stream (lib/bit-io.c, lib/misc.c:187-228, lib/arith.c) ---- */
typedef struct fi_bits
{
   FILE	   *file;
   uint8_t *buf;
   size_t   nbits, cap;		/* bits written, capacity in bytes */
   size_t   skip;		/* a stream in memory: leading bits that stand for the tail of what it will
				   be appended to (so that byte alignments fall where they do in the file) */
} fi_bits_t;

fi_bits_t *fi_bits_open (const char *filename);
fi_bits_t *fi_bits_open_mem (unsigned phase);
void	   fi_bits_free_mem (fi_bits_t *b);
void	   fi_bits_append (fi_bits_t *dst, const fi_bits_t *src);
void	   fi_bits_close (fi_bits_t *b);
void	   fi_put_bit (fi_bits_t *b, unsigned value);
void	   fi_put_bits (fi_bits_t *b, unsigned value, unsigned bits);
void	   fi_byte_align (fi_bits_t *b);
void	   fi_write_rice (fi_bits_t *b, unsigned value, unsigned rice_k);
void	   fi_write_bin_code (fi_bits_t *b, unsigned value, unsigned maxval);
void	   fi_encode_array (fi_bits_t *b, const unsigned *data, const unsigned *context,
			    const unsigned *c_symbols, unsigned n_context, unsigned n_data,
			    unsigned scaling);

/* 16-bit interval coder shared by the tree / matrix writers (lib/arith.h:98-121) */
typedef struct fi_ac
{
   uint16_t low, high;
   unsigned underflow;
   fi_bits_t *out;
} fi_ac_t;
void fi_ac_init (fi_ac_t *ac, fi_bits_t *out);
void fi_ac_rescale (fi_ac_t *ac);
void fi_ac_flush (fi_ac_t *ac);		/* low = high, rescale, byte align */

/* ---- the automaton as the writer sees it (codec/wfa.h:86-138) ---- */
typedef struct fi_wfainfo
{
   const char *basis_name, *title, *comment;
   unsigned    max_states, chroma_max_states;
   int	       color;
   unsigned    width, height, level;
   fi_rpf_t    rpf, dc_rpf, d_rpf, d_dc_rpf;
   unsigned    frames, fps, p_min_level, p_max_level, search_range;
   int	       half_pixel, B_as_past_ref;
   unsigned    smoothing;
   int	       nd_prediction;	/* coded with `--prediction': every frame carries its ND tree */
} fi_wfainfo_t;

typedef struct fi_wfa
{
   const fi_wfainfo_t *info;
   unsigned states, basis_states, root_state;
   const uint8_t  *level_of_state, *domain_type;
   const int16_t  (*tree) [2];
   const int16_t  (*into) [2][6];
   const float	  (*weight) [2][6];
   const int16_t  (*y_state) [2];
   const uint8_t  (*y_column) [2];
   /* predicted frames (NULL / 0 for an intra frame): codec/wfa.h:62-71,126,137 */
   int		   frame_type;		/* 0 intra, 1 predicted */
   const uint16_t (*x) [2], (*y) [2];	/* range coordinates (the motion tree asks for them) */
   const int8_t	  (*mv_type) [2], (*mv_fx) [2], (*mv_fy) [2], (*mv_bx) [2], (*mv_by) [2];
   const uint8_t  *delta_state;
} fi_wfa_t;

void fi_write_header (const fi_wfainfo_t *wi, fi_bits_t *out);
void fi_write_tables_init (void);
void fi_write_next_wfa (const fi_wfa_t *wfa, unsigned frame_number, int first,
			int normal_domains, int delta_domains, fi_bits_t *out);

#endif
