/*
 *  emu_runtime.cpp -- TEST INFRASTRUCTURE (see tests/emu/cuda_runtime.h): the fibre scheduler
 *  that runs one CUDA thread block at a time on one OS thread.
 *
 *  Every CUDA thread of the block is a fibre with its own stack; a fibre runs until it reaches
 *  a rendezvous (block barrier or warp collective) that is not complete yet, then the next
 *  fibre runs.  The last arrival completes the rendezvous and simply goes on.  Single OS
 *  thread: no data races, fully deterministic.  x86-64 only (hand-written context switch).
 */
#include "cuda_runtime.h"
#include <stdio.h>
#include <time.h>
#include <vector>

#if !defined(__x86_64__)
#error "the emulator's context switch is written for x86-64"
#endif

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

extern "C" void emu_switch (void **save_sp, void *new_sp);
asm (".text\n"
     ".globl emu_switch\n"
     ".type emu_switch,@function\n"
     "emu_switch:\n"
     "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
     "  movq %rsp, (%rdi)\n"
     "  movq %rsi, %rsp\n"
     "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
     "  ret\n"
     ".size emu_switch,.-emu_switch\n");

namespace {

enum { STACK_BYTES = 192 * 1024 };

struct Warp
{
   unsigned count, gen;
   unsigned slot [2][32];
};

enum { MAX_CLUSTER = 8 };

struct Block			/* the thread blocks of one cluster (usually one), run together */
{
   int	    n, alive, cur;	/* n: fibres of the whole cluster; per_cta threads each */
   int	    per_cta, ctas;
   unsigned first_block;	/* blockIdx.x of rank 0 */
   unsigned bar_count [MAX_CLUSTER], bar_gen [MAX_CLUSTER];
   int	    cta_alive [MAX_CLUSTER];
   unsigned cl_count, cl_gen;	/* barrier.cluster */
   std::vector<void *> sp;
   std::vector<char>   done;
   std::vector<Warp>   warps;
   void	   *main_sp;
   const std::function<void ()> *body;
};

Block	       B;
/* what each fibre waits for (0 runs, 1 block barrier, 2 cluster barrier, 3 warp collective), and
   the number of switches since a rendezvous last completed: a cycle through all fibres without
   progress is a deadlock */
std::vector<char> g_wait;
unsigned long	  g_idle;
/* scheduling order between rendezvous points (FB200_EMU_ORDER): 0 ascending thread index,
   1 descending, 2 pseudo-random.  A result that depends on it is a race between barriers. */
int	       g_order = -1;
unsigned       g_rand  = 12345u;
unsigned long long g_barriers, g_warp_ops;	/* rendezvous counters (whole process) */
char	      *g_stacks;
size_t	       g_stacks_n;
unsigned char *g_smem;		/* MAX_CLUSTER regions of SMEM_BYTES */
enum { SMEM_BYTES = 256 * 1024 };

inline int cur_rank (void) { return B.cur / B.per_cta; }

void
switch_to_next (void)
{
   const int from = B.cur;
   int	     next = from;

   if (++g_idle > (g_order == 2 ? 400ul : 4ul) * (unsigned long) B.n + 64)
   {
      int cnt [4] = {0, 0, 0, 0};
      for (int i = 0; i < B.n; i++)
	 if (!B.done [i])
	    cnt [(int) g_wait [i]]++;
      fprintf (stderr, "emu: deadlock -- %d threads alive: %d at a block barrier, %d at the cluster barrier, "
	       "%d in a warp collective, %d running\n", B.alive, cnt [1], cnt [2], cnt [3], cnt [0]);
      for (int r = 0; r < B.ctas; r++)
      {
	 int c [4] = {0, 0, 0, 0};
	 for (int i = r * B.per_cta; i < (r + 1) * B.per_cta; i++)
	    if (!B.done [i])
	       c [(int) g_wait [i]]++;
	 fprintf (stderr, "   rank %d: block barrier %d, cluster barrier %d, warp %d, running %d\n", r, c [1], c [2], c [3], c [0]);
      }
      abort ();
   }
   if (B.alive == 0)
   {
      B.cur = -1;
      emu_switch (&B.sp [from], B.main_sp);
      return;
   }
   do
   {
      if (g_order == 1)
	 next = next == 0 ? B.n - 1 : next - 1;
      else if (g_order == 2)
      {
	 g_rand = g_rand * 1664525u + 1013904223u;
	 next	= (int) ((g_rand >> 8) % (unsigned) B.n);
      }
      else
	 next = next + 1 == B.n ? 0 : next + 1;
   }
   while (B.done [next] || (g_order == 2 && next == from && B.alive > 1));
   if (next == from)
   {
      fprintf (stderr, "emu: deadlock -- thread %d waits for a rendezvous nobody else can reach\n", from);
      abort ();
   }
   B.cur       = next;
   threadIdx.x = (unsigned) (next % B.per_cta);
   blockIdx.x  = B.first_block + (unsigned) (next / B.per_cta);
   emu_switch (&B.sp [from], B.sp [next]);
}

void
fibre_main (void)
{
   (*B.body) ();
   const int r = cur_rank ();
   B.done [B.cur] = 1;
   g_idle = 0;
   B.alive--;
   B.cta_alive [r]--;
   if (B.cta_alive [r] && B.bar_count [r] == (unsigned) B.cta_alive [r])	/* the others wait at a barrier */
   {
      B.bar_count [r] = 0;
      B.bar_gen [r]++;
   }
   if (B.alive && B.cl_count == (unsigned) B.alive)
   {
      B.cl_count = 0;
      B.cl_gen++;
   }
   switch_to_next ();
   abort ();			/* a finished fibre is never resumed */
}

} /* namespace */

/* block barriers and warp collectives executed so far: the length of the dependent chain of a
   launch in units the GPU pays latency for (tools) */
extern "C" void
emu_counters (unsigned long long *barriers, unsigned long long *warp_ops)
{
   *barriers = g_barriers;
   *warp_ops = g_warp_ops;
}

void
emu_syncthreads (void)
{
   const int	  r   = cur_rank ();
   const unsigned gen = B.bar_gen [r];

   if (++B.bar_count [r] == (unsigned) B.cta_alive [r])
   {
      if (r == 0)
	 g_barriers++;
      B.bar_count [r] = 0;
      B.bar_gen [r]++;
      g_idle = 0;
      return;
   }
   g_wait [B.cur] = 1;
   while (B.bar_gen [r] == gen)
      switch_to_next ();
   g_wait [B.cur] = 0;
}

/* barrier.cluster.arrive + wait of every thread of the cluster */
void
emu_cluster_sync (void)
{
   const unsigned gen = B.cl_gen;

   if (++B.cl_count == (unsigned) B.alive)
   {
      B.cl_count = 0;
      B.cl_gen++;
      g_idle = 0;
      return;
   }
   g_wait [B.cur] = 2;
   while (B.cl_gen == gen)
      switch_to_next ();
   g_wait [B.cur] = 0;
}

unsigned emu_cluster_rank (void) { return (unsigned) cur_rank (); }
unsigned emu_cluster_size (void) { return (unsigned) B.ctas; }

/* mapa: the same shared-memory offset in the block of another rank */
void *
emu_map_shared_rank (const void *p, unsigned rank)
{
   const size_t off = (size_t) ((const unsigned char *) p - (g_smem + (size_t) cur_rank () * SMEM_BYTES));

   if (off >= SMEM_BYTES || rank >= (unsigned) B.ctas)
   {
      fprintf (stderr, "emu: map_shared_rank of an address outside dynamic shared memory\n");
      abort ();
   }
   return g_smem + (size_t) rank * SMEM_BYTES + off;
}

unsigned
emu_warp_exchange (unsigned value, int kind, int arg)
{
   const unsigned tid  = threadIdx.x;
   const unsigned lane = tid & 31u;
   Warp		 &w    = B.warps [(unsigned) B.cur >> 5];	/* per_cta is a multiple of 32 */
   const unsigned gen  = w.gen, buf = gen & 1u;
   const unsigned lanes = (tid | 31u) < (unsigned) B.per_cta ? 32u : (unsigned) B.per_cta - (tid & ~31u);

   w.slot [buf][lane] = value;
   if (++w.count == lanes)
   {
      if ((tid >> 5) == 0 && cur_rank () == 0)
	 g_warp_ops++;		/* warp 0's collectives: the resolution loops run in every warp alike */
      w.count = 0;
      w.gen++;
      g_idle = 0;
   }
   else
   {
      g_wait [B.cur] = 3;
      while (w.gen == gen)
	 switch_to_next ();
      g_wait [B.cur] = 0;
   }
   switch (kind)
   {
      case 0:
	 return w.slot [buf][(unsigned) arg & 31u];
      case 1:
	 return lane >= (unsigned) arg ? w.slot [buf][lane - (unsigned) arg] : value;
      case 2:
      {
	 unsigned m = 0;
	 for (unsigned l = 0; l < lanes; l++)
	    m |= (w.slot [buf][l] ? 1u : 0u) << l;
	 return m;
      }
      case 4:
	 return (lane ^ (unsigned) arg) < lanes ? w.slot [buf][lane ^ (unsigned) arg] : value;
      default:
	 return 0;
   }
}

unsigned char *
emu_dyn_smem (void)
{
   return g_smem + (size_t) cur_rank () * SMEM_BYTES;
}

long long
clock64 (void)
{
   struct timespec ts;
   clock_gettime (CLOCK_MONOTONIC, &ts);
   return (long long) ts.tv_sec * 1000000000LL + ts.tv_nsec;
}

void
emu_launch_cluster (emu_dim3 grid, emu_dim3 block, unsigned cluster, size_t smem, const std::function<void ()> &body)
{
   const int per_cta = (int) block.x;
   const int n	     = per_cta * (int) cluster;

   {
      const char *o = getenv ("FB200_EMU_ORDER");	/* read per launch: tests switch it */
      g_order = o ? atoi (o) : 0;
   }
   if (!g_smem && posix_memalign ((void **) &g_smem, 256, (size_t) MAX_CLUSTER * SMEM_BYTES))
      abort ();
   if (smem > SMEM_BYTES || block.y != 1 || block.z != 1 || cluster < 1 || cluster > MAX_CLUSTER
       || grid.x % cluster || (cluster > 1 && per_cta % 32))
   {
      fprintf (stderr, "emu: launch shape not supported\n");
      abort ();
   }
   if (g_stacks_n < (size_t) n)
   {
      free (g_stacks);
      if (posix_memalign ((void **) &g_stacks, 4096, (size_t) n * STACK_BYTES))
	 abort ();
      g_stacks_n = (size_t) n;
   }
   gridDim  = grid;
   blockDim = block;
   for (unsigned by = 0; by < grid.y; by++)
   for (unsigned bx = 0; bx < grid.x; bx += cluster)
   {
      for (unsigned r = 0; r < cluster; r++)
	 memset (g_smem + (size_t) r * SMEM_BYTES, 0xcd, smem);	/* shared memory starts undefined */
      B.n = B.alive = n;
      B.per_cta	    = per_cta;
      B.ctas	    = (int) cluster;
      B.first_block = bx;
      B.cl_count = B.cl_gen = 0;
      for (unsigned r = 0; r < MAX_CLUSTER; r++)
      {
	 B.bar_count [r] = B.bar_gen [r] = 0;
	 B.cta_alive [r] = per_cta;
      }
      g_wait.assign ((size_t) n, 0);
      g_idle = 0;
      B.sp.assign ((size_t) n, NULL);
      B.done.assign ((size_t) n, 0);
      B.warps.assign ((size_t) (n + 31) / 32, Warp ());
      B.body = &body;
      for (int t = 0; t < n; t++)
      {
	 /* initial frame: six callee-saved registers, the entry point, a dummy return slot;
	    after the ret the stack pointer is 8 mod 16 as at any function entry */
	 void **top = (void **) (g_stacks + (size_t) (t + 1) * STACK_BYTES);
	 void **sp  = top - 8;
	 for (int i = 0; i < 6; i++)
	    sp [i] = NULL;
	 sp [6] = (void *) fibre_main;
	 sp [7] = NULL;
	 B.sp [t] = sp;
      }
      B.cur	  = g_order == 1 ? n - 1 : 0;
      threadIdx.x = (unsigned) (B.cur % per_cta);
      blockIdx	  = emu_dim3 (bx + (unsigned) (B.cur / per_cta), by, 0);
      emu_switch (&B.main_sp, B.sp [B.cur]);
   }
}

void
emu_launch (emu_dim3 grid, emu_dim3 block, size_t smem, const std::function<void ()> &body)
{
   emu_launch_cluster (grid, block, 1, smem, body);
}
