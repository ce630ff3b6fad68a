/*
 *  pnm_input.c -- raw PGM / PPM reader and the coder's pixel format.
 *  What must match the reference (lib/image.c:262-388): P5 / P6 only, sizes >= 32 and even,
 *  grey sample g -> (g - 128) * 16, colour samples -> Y, Cb, Cr with the reference's
 *  double precision matrix, scaled by 16 and truncated to short.  Always 4:4:4.
 */
#include <errno.h>
#include <stdlib.h>
#include <string.h>

#include "fi_internal.h"

static void
skip_space_and_comments (FILE *f)
{
   int c;

   for (;;)
   {
      c = fgetc (f);
      if (c == '#')
      {
	 while ((c = fgetc (f)) != '\n' && c != EOF)
	    ;
      }
      else if (c != ' ' && c != '\t' && c != '\r' && c != '\n')
      {
	 if (c != EOF)
	    ungetc (c, f);
	 return;
      }
   }
}

/* an error while a file is open: the handle does not outlive the message (fi_error() does not return) */
#define PNM_ERROR(f, ...) do { if ((f) && (f) != stdin) fclose (f); fi_error (__VA_ARGS__); } while (0)

static int
read_number (FILE *f)
{
   int value = 0, c, digits = 0;

   skip_space_and_comments (f);
   while ((c = fgetc (f)) >= '0' && c <= '9')
   {
      value = 10 * value + (c - '0');
      digits++;
   }
   if (c != EOF)
      ungetc (c, f);
   if (!digits)
      PNM_ERROR (f, "Format error: can't read PNM header.");
   return value;
}

static FILE *
open_pnm (const char *name, unsigned *width, unsigned *height, int *color)
{
   FILE *f = open_file (name, "FIASCO_IMAGES", READ_ACCESS);
   int	 m0, m1, v;

   if (f == NULL)
      fi_file_error (name ? name : "stdin");
   m0 = fgetc (f);
   m1 = fgetc (f);
   if (m0 == 'P' && m1 == '5')
      *color = 0;
   else if (m0 == 'P' && m1 == '6')
      *color = 1;
   else
      PNM_ERROR (f, "%s: image format '%c%c' not supported.", name ? name : "stdin", m0, m1);
   v = read_number (f);
   if (v < 32)
      PNM_ERROR (f, "Width of image `%s' has to be at least 32 pixels.", name ? name : "stdin");
   *width = (unsigned) v;
   v = read_number (f);
   if (v < 32)
      PNM_ERROR (f, "Height of image `%s' has to be at least 32 pixels.", name ? name : "stdin");
   *height = (unsigned) v;
   (void) read_number (f);		/* maxval */
   if (fgetc (f) == EOF)		/* the single white space before the raster */
      PNM_ERROR (f, "%s: EOF reached, input seems to be truncated!", name ? name : "stdin");
   return f;
}

void
fi_read_pnm_header (const char *name, unsigned *width, unsigned *height, int *color)
{
   FILE *f = open_pnm (name, width, height, color);

   if (f != stdin)
      fclose (f);
}

fi_image_t *
fi_read_image (const char *name)
{
   unsigned    width, height, n, i;
   int	       color;
   FILE	      *f   = open_pnm (name, &width, &height, &color);
   fi_image_t *img;

   if ((width & 1) || (height & 1))
      PNM_ERROR (f, "Width and height of images must be even numbers.");
   img	       = fiasco_calloc (1, sizeof (fi_image_t));
   img->width  = width;
   img->height = height;
   img->color  = color;
   n	       = width * height;
   for (i = 0; i < (color ? 3u : 1u); i++)	/* (every sample is written below: no need for zeroed memory) */
      if (!(img->pixels [i] = malloc ((size_t) n * sizeof (int16_t))))
	 PNM_ERROR (f, "Out of memory!");
   {
      /* the raster in one read; a short read is the reference's I/O error (lib/image.c:357-388) */
      const size_t   bytes = (size_t) n * (color ? 3u : 1u);
      unsigned char *raw   = malloc (bytes);

      if (!raw)
	 fi_error ("Out of memory!");
      if (fread (raw, 1, bytes, f) != bytes)
      {
	 const int why = errno;		/* (closing the file must not change what the message says) */

	 free (raw);
	 if (f != stdin)
	    fclose (f);
	 errno = why;
	 fi_file_error (name ? name : "stdin");
      }
      if (!color)
	 for (i = 0; i < n; i++)
	    img->pixels [0][i] = (int16_t) (((int) raw [i] - 128) * 16);
      else
	 for (i = 0; i < n; i++)
	 {
	    const int r = raw [3 * i], g = raw [3 * i + 1], b = raw [3 * i + 2];

	    img->pixels [0][i] = (int16_t) ((+0.2989 * r + 0.5866 * g + 0.1145 * b - 128) * 16);
	    img->pixels [1][i] = (int16_t) ((-0.1687 * r - 0.3312 * g + 0.5000 * b) * 16);
	    img->pixels [2][i] = (int16_t) ((+0.5000 * r - 0.4183 * g - 0.0816 * b) * 16);
	 }
      free (raw);
   }
   if (f != stdin)
      fclose (f);
   return img;
}

void
fi_free_image (fi_image_t *image)
{
   int i;

   if (!image)
      return;
   for (i = 0; i < 3; i++)
      free (image->pixels [i]);
   free (image);
}
