/*
 *  misc.c -- the two non-fiasco.h symbols the reference command line tools link
 *  (SURVEY.md 8b): fiasco_calloc (lib/misc.c:51) and open_file (lib/bit-io.c:48), plus
 *  fiasco_free.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include "fi_internal.h"

#ifndef FIASCO_SHARE
#define FIASCO_SHARE "/usr/local/share/fiasco/"
#endif

void *
fiasco_calloc (size_t n, size_t size)
{
   void *ptr;

   if (n <= 0 || size <= 0)
      fi_error ("Can't allocate memory for %d items of size %d", (int) n, (int) size);
   ptr = calloc (n, size);
   if (ptr == NULL)
      fi_error ("Out of memory!");
   return ptr;
}

void
fiasco_free (void *ptr)
{
   if (ptr != NULL)
      free (ptr);
   else
      fi_warning ("Can't free memory block <NULL>.");
}

static FILE *
try_dir (const char *dir, const char *filename, const char *mode)
{
   size_t len  = strlen (dir);
   char	 *path = malloc (len + strlen (filename) + 2);
   FILE	 *fp;

   strcpy (path, dir);
   if (len == 0 || path [len - 1] != '/')
      strcat (path, "/");
   strcat (path, filename);
   fp = fopen (path, mode);
   free (path);
   return fp;
}

/*
 *  Search order of the reference (lib/bit-io.c:48-143): "-"/NULL is stdin/stdout; a
 *  readable file in the current directory; a name with a '/' is written as given;
 *  otherwise every directory of $env_var (separators " ;:,", default "."), finally the
 *  share directory.
 */
FILE *
open_file (const char *filename, const char *env_var, openmode_e mode)
{
   const char *fmode = mode == READ_ACCESS ? "r" : "w";
   const char *env   = NULL;
   char	      *dirs, *dir;
   FILE	      *fp = NULL;

   if (filename == NULL || strcmp (filename, "-") == 0)
      return mode == READ_ACCESS ? stdin : stdout;
   if (mode == READ_ACCESS && (fp = fopen (filename, fmode)))
      return fp;
   if (mode == WRITE_ACCESS && strchr (filename, '/'))
      return fopen (filename, fmode);
   if (env_var != NULL)
      env = getenv (env_var);
   dirs = strdup (env ? env : ".");
   for (dir = strtok (dirs, " ;:,"); dir && !fp; dir = strtok (NULL, " ;:,"))
      fp = try_dir (dir, filename, fmode);
   if (fp == NULL)
      fp = try_dir (FIASCO_SHARE, filename, fmode);
   free (dirs);
   return fp;
}

/*****************************************************************************
		host threads for the per-frame work around the launches
*****************************************************************************/

typedef struct pfor
{
   size_t	   n;
   volatile size_t next;
   volatile int	   failed;
   char		   err [2048];	/* message of the first call that failed */
   void		 (*fn) (size_t, void *);
   void		  *ctx;
   pthread_mutex_t lock;
} pfor_t;

static void *
pfor_worker (void *arg)
{
   pfor_t *p = arg;

   for (;;)
   {
      size_t i;

      pthread_mutex_lock (&p->lock);
      i = p->next++;
      pthread_mutex_unlock (&p->lock);
      if (i >= p->n || p->failed)
	 return NULL;
      /* fi_error () jumps to the thread's own fi_env */
      fi_try
      {
	 p->fn (i, p->ctx);
      }
      fi_catch
      {
	 pthread_mutex_lock (&p->lock);
	 if (!p->failed)
	    snprintf (p->err, sizeof p->err, "%s", fiasco_get_error_message ());
	 p->failed = 1;
	 pthread_mutex_unlock (&p->lock);
	 return NULL;
      }
   }
}

int
fi_parallel_for (size_t n, void (*fn) (size_t i, void *ctx), void *ctx)
{
   pfor_t      p;
   pthread_t   th [16];
   unsigned    threads = 0, started = 0;
   const char *e = getenv ("FIASCO_HOST_THREADS");
   jmp_buf     caller;

   if (e && atoi (e) > 0)
      threads = (unsigned) atoi (e);
   else
   {
      cpu_set_t set;

      threads = sched_getaffinity (0, sizeof set, &set) == 0 ? (unsigned) CPU_COUNT (&set) : 1;
   }
   if (threads > 16)
      threads = 16;
   if (threads > n)
      threads = (unsigned) n;
   memset (&p, 0, sizeof p);
   p.n	 = n;
   p.fn	 = fn;
   p.ctx = ctx;
   pthread_mutex_init (&p.lock, NULL);
   memcpy (caller, fi_env, sizeof caller);	/* the calling thread takes part with a try of its own */
   for (; started + 1 < threads; started++)
      if (pthread_create (&th [started], NULL, pfor_worker, &p))
	 break;
   pfor_worker (&p);
   for (unsigned t = 0; t < started; t++)
      pthread_join (th [t], NULL);
   memcpy (fi_env, caller, sizeof caller);
   pthread_mutex_destroy (&p.lock);
   if (p.failed)
      fi_set_error ("%s", p.err);	/* the message lives per thread: hand it to the caller's */
   return p.failed;
}
