/*
 *  error.c -- error handling and messages of the host library.
 *  Behaviour follows the reference's lib/error.c: fi_error() stores the text and jumps
 *  back to the public entry point (which then returns 0), fiasco_get_error_message()
 *  returns the stored text, messages go to stderr gated by the verbosity level.
 */
#include <errno.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "fi_internal.h"

/* per thread: the GPU worker threads of fiasco_coder() run host code that reports errors too */
__thread jmp_buf fi_env;

static fiasco_verbosity_e verboselevel = FIASCO_SOME_VERBOSITY;
static __thread char	  error_message [2048];

static void
store (const char *format, va_list args)
{
   vsnprintf (error_message, sizeof error_message, format, args);
}

void
fi_set_error (const char *format, ...)
{
   va_list args;

   va_start (args, format);
   store (format, args);
   va_end (args);
}

void
fi_error (const char *format, ...)
{
   va_list args;

   va_start (args, format);
   store (format, args);
   va_end (args);
   longjmp (fi_env, 1);
}

/* jump with the message that is already stored */
void
fi_rethrow (void)
{
   longjmp (fi_env, 1);
}

const char *
fi_system_error (void)
{
   return strerror (errno);
}

void
fi_file_error (const char *filename)
{
   fi_error ("File `%s': I/O Error - %s.", filename, fi_system_error ());
}

const char *
fiasco_get_error_message (void)
{
   return error_message;
}

void
fi_warning (const char *format, ...)
{
   va_list args;

   if (verboselevel == FIASCO_NO_VERBOSITY)
      return;
   va_start (args, format);
   fprintf (stderr, "Warning: ");
   vfprintf (stderr, format, args);
   fputc ('\n', stderr);
   va_end (args);
}

void
fi_message (const char *format, ...)
{
   va_list args;

   if (verboselevel == FIASCO_NO_VERBOSITY)
      return;
   va_start (args, format);
   vfprintf (stderr, format, args);
   fputc ('\n', stderr);
   va_end (args);
}

/* lib/error.c:279: no newline, flushed (the progress meter) */
void
fi_info (const char *format, ...)
{
   va_list args;

   if (verboselevel == FIASCO_NO_VERBOSITY)
      return;
   va_start (args, format);
   vfprintf (stderr, format, args);
   fflush (stderr);
   va_end (args);
}

void
fi_debug_message (const char *format, ...)
{
   va_list args;

   if (verboselevel < FIASCO_ULTIMATE_VERBOSITY)
      return;
   va_start (args, format);
   fprintf (stderr, "*** ");
   vfprintf (stderr, format, args);
   fputc ('\n', stderr);
   va_end (args);
}

void
fiasco_set_verbosity (fiasco_verbosity_e level)
{
   verboselevel = level;
}

fiasco_verbosity_e
fiasco_get_verbosity (void)
{
   return verboselevel;
}
