/*
 *  misc.c -- the two non-fiasco.h symbols the reference command line tools link
 *  (SURVEY.md 8b): fiasco_calloc (lib/misc.c:51) and open_file (lib/bit-io.c:48), plus
 *  fiasco_free.
 */
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
