/*
 *  oracle_enc.c -- command line front end of the CPU restatement (TEST INFRASTRUCTURE).
 *
 *  usage: oracle_enc in.pnm quality optimize [dump.txt [trace.txt]]
 *  Reads a raw PGM/PPM (maxval 255), encodes it with fo_encode() and prints the WFA in
 *  the grammar of oracle/wfadump.c, so that
 *      diff <(oracle/_ref/wfadump ref.fco | grep '^[se] ') <(oracle_enc ... | grep '^[se] ')
 *  is the parity check against the reference.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fiasco_oracle.h"

static int
read_int (FILE *f)
{
   int c, v = 0;

   for (;;)
   {
      c = fgetc (f);
      if (c == '#')
	 while ((c = fgetc (f)) != '\n' && c != EOF)
	    ;
      else if (c >= '0' && c <= '9')
	 break;
      else if (c == EOF)
	 return -1;
   }
   for (; c >= '0' && c <= '9'; c = fgetc (f))
      v = 10 * v + c - '0';
   return v;
}

int
main (int argc, char **argv)
{
   FILE	      *in, *dump = stdout, *trace = NULL;
   char	       magic [3] = {0};
   int	       width, height, color, n;
   uint8_t    *raw;
   int16_t    *planes [3] = {NULL, NULL, NULL};
   fo_params_t p;
   fo_wfa_t   *wfa = malloc (sizeof (fo_wfa_t));
   fo_stats_t  st;
   char	       err [256] = "";
   clock_t     t0;

   if (argc < 4)
   {
      fprintf (stderr, "usage: %s in.pnm quality optimize [dump [trace]]\n", argv [0]);
      return 2;
   }
   if (!(in = fopen (argv [1], "rb")) || fread (magic, 1, 2, in) != 2)
   {
      fprintf (stderr, "can't read %s\n", argv [1]);
      return 1;
   }
   color  = magic [1] == '6';
   width  = read_int (in);
   height = read_int (in);
   (void) read_int (in);
   n   = width * height;
   raw = malloc ((size_t) n * (color ? 3 : 1));
   if (fread (raw, color ? 3 : 1, (size_t) n, in) != (size_t) n)
   {
      fprintf (stderr, "truncated input\n");
      return 1;
   }
   fclose (in);
   planes [0] = malloc (sizeof (int16_t) * (size_t) n);
   if (color)
   {
      planes [1] = malloc (sizeof (int16_t) * (size_t) n);
      planes [2] = malloc (sizeof (int16_t) * (size_t) n);
      fo_rgb_to_planes (raw, (size_t) n, planes [0], planes [1], planes [2]);
   }
   else
      fo_grey_to_plane (raw, (size_t) n, planes [0]);

   fo_default_params (&p, width, height, color, (float) atof (argv [2]), atoi (argv [3]));
   if (argc > 4 && strcmp (argv [4], "-"))
      dump = fopen (argv [4], "w");
   if (argc > 5)
      trace = fopen (argv [5], "w");
   t0 = clock ();
   if (fo_encode (&p, (const int16_t *const *) planes, wfa, &st, trace, err, sizeof err))
   {
      fprintf (stderr, "oracle_enc: %s\n", err);
      return 1;
   }
   fprintf (stderr, "oracle: %.3f s, %u states, subdivide %llu, mp %llu (avg D %.1f), "
	    "pass1 %llu, pass2 %llu, ortho %llu, accepted %llu, leaf dots %llu, "
	    "ipss lookups %llu, blocks %llu, ip bytes %llu\n",
	    (double) (clock () - t0) / CLOCKS_PER_SEC, wfa->states,
	    (unsigned long long) st.subdivide_calls, (unsigned long long) st.mp_calls,
	    st.mp_calls ? (double) st.mp_domains / (double) st.mp_calls : 0.0,
	    (unsigned long long) st.pass1, (unsigned long long) st.pass2,
	    (unsigned long long) st.ortho_steps, (unsigned long long) st.accepted,
	    (unsigned long long) st.leaf_dots, (unsigned long long) st.ipss_lookups,
	    (unsigned long long) st.blocks, (unsigned long long) st.ip_bytes);
   fo_dump_wfa (wfa, &p, dump);
   if (trace)
      fclose (trace);
   return 0;
}
