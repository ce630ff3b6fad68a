/*
 *  bitstream.c -- bit-granular output and the entropy coding primitives of the .fco
 *  stream: MSB-first bits, Rice code, adjusted binary code, the 16-bit interval coder
 *  and the adaptive multi-context array coder.
 *
 *  The byte sequence produced must equal the reference's (lib/bit-io.c:260-330,
 *  lib/misc.c:187-228, lib/arith.h:98-121, lib/arith.c:197-306); the implementation
 *  keeps the whole stream in memory and writes it when the file is closed.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fi_internal.h"

fi_bits_t *
fi_bits_open (const char *filename)
{
   fi_bits_t *b = fiasco_calloc (1, sizeof (fi_bits_t));

   b->file = open_file (filename, "FIASCO_DATA", WRITE_ACCESS);
   if (b->file == NULL)
   {
      free (b);
      return NULL;
   }
   b->cap = 1 << 16;
   b->buf = fiasco_calloc (b->cap, 1);
   return b;
}

/* a stream that is not tied to a file: frames are coded side by side into streams of their own
   and appended to the output in coding order (fi_bits_append) */
fi_bits_t *
fi_bits_open_mem (unsigned phase)
{
   fi_bits_t *b = fiasco_calloc (1, sizeof (fi_bits_t));

   b->file  = NULL;
   b->cap   = 1 << 14;
   b->buf   = fiasco_calloc (b->cap, 1);
   b->nbits = b->skip = phase & 7;
   return b;
}

void
fi_bits_free_mem (fi_bits_t *b)
{
   if (b)
   {
      free (b->buf);
      free (b);
   }
}

void
fi_bits_append (fi_bits_t *dst, const fi_bits_t *src)
{
   if ((dst->nbits & 7) == 0 && src->skip == 0)
   {
      const size_t at = dst->nbits >> 3, bytes = (src->nbits + 7) >> 3;

      if (at + bytes + 1 > dst->cap)
      {
	 size_t cap = dst->cap;

	 while (at + bytes + 1 > cap)
	    cap *= 2;
	 dst->buf = realloc (dst->buf, cap);
	 if (!dst->buf)
	    fi_error ("Out of memory!");
	 memset (dst->buf + dst->cap, 0, cap - dst->cap);
	 dst->cap = cap;
      }
      memcpy (dst->buf + at, src->buf, bytes);
      dst->nbits += src->nbits;
   }
   else
      for (size_t i = src->skip; i < src->nbits; i++)
	 fi_put_bit (dst, (src->buf [i >> 3] >> (7 - (i & 7))) & 1u);
}

void
fi_put_bit (fi_bits_t *b, unsigned value)
{
   const size_t byte = b->nbits >> 3;

   if (byte >= b->cap)
   {
      b->buf = realloc (b->buf, b->cap * 2);
      if (!b->buf)
	 fi_error ("Out of memory!");
      memset (b->buf + b->cap, 0, b->cap);
      b->cap *= 2;
   }
   if (value)
      b->buf [byte] |= (uint8_t) (0x80u >> (b->nbits & 7));
   b->nbits++;
}

void
fi_put_bits (fi_bits_t *b, unsigned value, unsigned bits)
{
   while (bits--)
      fi_put_bit (b, (value >> bits) & 1u);
}

void
fi_byte_align (fi_bits_t *b)
{
   while (b->nbits & 7)
      fi_put_bit (b, 0);
}

void
fi_bits_close (fi_bits_t *b)
{
   /* the reference always emits the byte its cursor stands on, i.e. at least one byte
      (lib/bit-io.c:308-318) */
   size_t bytes = (b->nbits + 7) >> 3;

   if (bytes == 0)
      bytes = 1;
   if (fwrite (b->buf, 1, bytes, b->file) != bytes)
      fi_error ("Can't write remaining %d bytes of bitfile!", (int) bytes);
   if (b->file != stdout)
      fclose (b->file);
   else
      fflush (stdout);
   free (b->buf);
   free (b);
}

void
fi_write_rice (fi_bits_t *b, unsigned value, unsigned rice_k)
{
   unsigned unary;

   for (unary = value >> rice_k; unary; unary--)
      fi_put_bit (b, 1);
   fi_put_bit (b, 0);
   fi_put_bits (b, value & ((1u << rice_k) - 1), rice_k);
}

void
fi_write_bin_code (fi_bits_t *b, unsigned value, unsigned maxval)
{
   unsigned k = 0, r;

   while ((2u << k) <= maxval + 1)	/* k = floor (log2 (maxval + 1)) */
      k++;
   r = (maxval + 1) - (1u << k);
   if (value < maxval + 1 - 2 * r)
      fi_put_bits (b, value, k);
   else
      fi_put_bits (b, value + maxval + 1 - 2 * r, k + 1);
}

/* ---- interval coder: 16-bit low/high with pending underflow bits ---- */

void
fi_ac_init (fi_ac_t *ac, fi_bits_t *out)
{
   ac->low	 = 0x0000;
   ac->high	 = 0xffff;
   ac->underflow = 0;
   ac->out	 = out;
}

void
fi_ac_rescale (fi_ac_t *ac)
{
   for (;;)
   {
      if (ac->high < 0x8000)
      {
	 fi_put_bit (ac->out, 0);
	 for (; ac->underflow; ac->underflow--)
	    fi_put_bit (ac->out, 1);
      }
      else if (ac->low >= 0x8000)
      {
	 fi_put_bit (ac->out, 1);
	 for (; ac->underflow; ac->underflow--)
	    fi_put_bit (ac->out, 0);
      }
      else if (ac->high < 0xc000 && ac->low >= 0x4000)
      {
	 ac->underflow++;
	 ac->high |= 0x4000;
	 ac->low  &= 0x3fff;
      }
      else
	 break;
      ac->high = (uint16_t) ((ac->high << 1) | 1);
      ac->low  = (uint16_t) (ac->low << 1);
   }
}

void
fi_ac_flush (fi_ac_t *ac)
{
   ac->low = ac->high;
   fi_ac_rescale (ac);
   fi_byte_align (ac->out);
}

/*
 *  Adaptive arithmetic coding of an array with one frequency table per context, every
 *  table starting uniform, halved when its total exceeds 'scaling' (lib/arith.c:197-306).
 *  The reference keeps the cumulative counts in 16-bit words.
 */
void
fi_encode_array (fi_bits_t *b, const unsigned *data, const unsigned *context,
		 const unsigned *c_symbols, unsigned n_context, unsigned n_data,
		 unsigned scaling)
{
   uint16_t **totals;
   fi_ac_t    ac;
   unsigned   c, i, n;

   if (!n_context)
      n_context = 1;
   totals = fiasco_calloc (n_context, sizeof (uint16_t *));
   for (c = 0; c < n_context; c++)
   {
      totals [c] = fiasco_calloc (c_symbols [c] + 1, sizeof (uint16_t));
      for (i = 0; i < c_symbols [c]; i++)
	 totals [c][i + 1] = (uint16_t) (totals [c][i] + 1);
   }
   fi_ac_init (&ac, b);
   for (n = 0; n < n_data; n++)
   {
      /* The reference takes the symbol as an int (lib/arith.c:251): the code of a weight that rounds
	 to zero, RPF_ZERO = -1 -- which only the DC weight of a nondeterministic prediction can be
	 (output/nd.c:222, no zero check there) -- makes it read the 16-bit word in front of its table:
	 the upper end of glibc's chunk size field, zero for every chunk below 2^48 bytes.  The same
	 value is used here; the update below then starts at entry 0, as it does there. */
      const int d = (int) data [n];
      unsigned	range;
      uint16_t	scale, low_count, high_count;

      c		 = n_context > 1 ? context [n] : 0;
      scale	 = totals [c][c_symbols [c]];
      low_count	 = d < 0 ? 0 : totals [c][d];
      high_count = totals [c][d + 1];
      range	 = (unsigned) (ac.high - ac.low) + 1;
      ac.high	 = (uint16_t) (ac.low + (uint16_t) ((range * high_count) / scale - 1));
      ac.low	 = (uint16_t) (ac.low + (uint16_t) ((range * low_count) / scale));
      fi_ac_rescale (&ac);
      for (i = (unsigned) (d + 1); i < c_symbols [c] + 1; i++)
	 totals [c][i]++;
      if (totals [c][c_symbols [c]] > scaling)
	 for (i = 1; i < c_symbols [c] + 1; i++)
	 {
	    totals [c][i] >>= 1;
	    if (totals [c][i] <= totals [c][i - 1])
	       totals [c][i] = (uint16_t) (totals [c][i - 1] + 1);
	 }
   }
   fi_ac_flush (&ac);
   for (c = 0; c < n_context; c++)
      free (totals [c]);
   free (totals);
}

/* ---- reduced precision format (lib/rpf.c:59-111, :171-222), host copy ---- */

fi_rpf_t
fi_make_rpf (unsigned mantissa, int range_e)
{
   fi_rpf_t r;

   if (mantissa < 2 || mantissa > 8)
      mantissa = 2;		/* the reference maps both out-of-range cases to 2 */
   r.mantissa_bits = mantissa;
   r.range_e	   = range_e;
   r.range	   = range_e == 0 ? 0.75f : range_e == 2 ? 1.5f : range_e == 3 ? 2.0f : 1.0f;
   return r;
}

int
fi_rtob (float f, const fi_rpf_t *rpf)
{
   uint32_t u, mantissa;
   int	    exponent, sign;

   f /= rpf->range;
   memcpy (&u, &f, 4);
   mantissa = ((u & 0x7fffffu) >> 1) | (1u << 22);
   exponent = (int) ((u >> 23) & 0xffu) - 126;
   sign	    = (int) (u >> 31);
   if (exponent > 0)
      mantissa <<= (exponent & 31);
   else
      mantissa >>= ((-exponent) & 31);
   mantissa >>= (23 - rpf->mantissa_bits - 1);
   mantissa   = (mantissa + 1) >> 1;
   if (mantissa == 0)
      return -1;
   if (mantissa >= (1u << rpf->mantissa_bits))
      return sign;
   return (int) (((mantissa & ((1u << rpf->mantissa_bits) - 1)) << 1) | (unsigned) sign);
}
