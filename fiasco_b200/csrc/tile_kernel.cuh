/*
 *  tile_kernel.cuh -- shared declarations of the persistent tile kernel and its host
 *  launcher (internal; the public boundary is include/fiasco_b200.h).
 */
#ifndef FB200_TILE_KERNEL_CUH
#define FB200_TILE_KERNEL_CUH

#include <stdint.h>
#include <cuda_runtime.h>
#include "fiasco_b200.h"

/*
 *  Kernel launch and dynamic shared memory, spelled so that the CPU-only test suite can compile
 *  these sources as plain C++ against tests/emu/cuda_runtime.h (test infrastructure: a fibre
 *  per CUDA thread; never built into the product).
 */
#ifdef FB200_EMU
#define FB_LAUNCH(kernel, grid, block, smem, stream, ...) \
   emu_launch (grid, block, smem, [&] () { kernel (__VA_ARGS__); })
#define FB_LAUNCH_CLUSTER(kernel, grid, block, cluster, smem, stream, ...) \
   emu_launch_cluster (grid, block, cluster, smem, [&] () { kernel (__VA_ARGS__); })
#define FB_DYN_SMEM(type, name) type *name = (type *) emu_dyn_smem ()
#else
#define FB_LAUNCH(kernel, grid, block, smem, stream, ...) \
   kernel<<<grid, block, smem, stream>>> (__VA_ARGS__)
#define FB_DYN_SMEM(type, name) extern __shared__ __align__ (16) type name []
#endif

#define FB_MAXEDGES  5
#define FB_NO_EDGE   (-1)
#define FB_RANGE     (-1)
#define FB_MAXCOSTS  1e20f		/* codec/coder.c:53 */
#define FB_MAXDEPTH  (FB200_MAXLEVEL + 2)
#define FB_IMG_STRIDE 64		/* floats per state image row (63 used, il <= 5) */

/* layout of the probability-model blob (int16 units) */
#define MB_COUNT      0			/* rle count[6]	  (domain-pool.c:623) */
#define MB_TOTAL      6			/* rle total (u16) */
#define MB_N	      7			/* rle n (u16) */
#define MB_MAXDOM     8			/* rle max_domains (u16) */
#define MB_YINDEX     9			/* rle y_index (u16) */
#define MB_D0N	      10		/* DC qac model: n */
#define MB_D0INDEX    11		/* DC qac model: index[0] */
#define MB_D0YINDEX   12		/* DC qac model: y_index (u16) */
#define MB_TOTALS     16		/* aac totals[1 + levels] */
#define MB_TOTALS_MAX 24
#define MB_COUNTS     (MB_TOTALS + MB_TOTALS_MAX + 1) /* aac counts, 1 pad slot before */

struct DevParams
{
   int	 width, height, level, bands;
   int	 lc_min, lc_max, il;	/* range levels, images level */
   int	 lmin;			/* lowest level with tables = min (lc_min, il) */
   int	 nlev;			/* lc_max - lmin + 1 */
   int	 tn;			/* nodes of the per-block product tree: 2^nlev - 1 */
   int	 max_elements, max_domains, chroma_max_states;
   float price, chroma_decrease;
   int	 rpf_m, dc_m;
   float rpf_range, dc_range;
   int	 second_domain_block;
   int	 s_cap;			/* state capacity */
   int	 coeff_min_level;	/* context base of the aac model (coder.c:731) */
   int	 aac_dc_size, aac_lvl_size, blob_len;
   int	 trace_cap;
   unsigned sm_off [20];	/* shared-memory layout of the block (filled by the launcher) */
   int	 n_slots;		/* workspaces (img, T, SS, ...); tiles beyond that share them */
   int	*slot_flags;		/* [n_slots] 0 = free */
   int	 big;			/* large state capacity, so that more tiles fit on an SM: bit 0 =
				   Gram rows in global memory, bit 1 = model snapshots in global */
   /* predicted frames (codec/prediction.c, codec/mwfa.c): 0 = intra kernel */
   int	 motion;		/* frame type: 0 intra, 1 P frame, 2 B frame */
   int	 p_min, p_max;		/* levels of the motion compensated ranges (coder.c:284-290) */
   int	 sr;			/* search range: vectors in [-sr, sr) */
   int	 blob_half;		/* shorts of one model set; blob = normal set, then delta set */
   int	 n_frames;		/* DFS activation records (nested delta pass included) */
   const int *tile_lc_min;	/* [tiles] lc_min_level a tile starts with (0: lc_min), or NULL */
   const int *tile_ready;	/* [tiles] or NULL: a block starts when tile_ready [tile] == ready_epoch -- the
				   pixels of the later tiles of a batch travel while the first ones are coded */
   int	 ready_epoch;
   int	 cluster;		/* thread blocks per stream (filled by the launcher): rank 0 walks the
				   recursion, the others serve it (tile_kernel.cu, "cluster per stream") */
};
#define FB_MAXCLUSTER 8

/* all transitions of one state in one 64-byte line: what the inner-product kernels gather */
struct __align__ (64) Trans
{
   short child [2];		/* wfa->tree [s][label] */
   short into [2][FB_MAXEDGES];	/* -1 terminated unless all 5 are used */
   float w [2][FB_MAXEDGES];
};

/* per-tile device workspace (all pointers device memory) */
struct TileWs
{
   const int16_t *pix;		/* [bands][width*height] */
   float   *img;		/* [s_cap][64]		  state images, levels 0..il */
   float   *T;			/* [tn][s_cap]		  range x state products */
   float   *SS;			/* [nlev][s_cap][s_cap]	  state x state products */
   float   *diag;		/* [nlev][s_cap]	  <s,s> */
   Trans   *trans;		/* [s_cap]		  packed transitions */
   float   *Gglob;		/* [FB_MAXCLUSTER][max_elements-1][s_cap+1] Gram-Schmidt rows when not in
				   smem, one set per block of the cluster */
   /* automaton */
   float   *final_d;		/* [s_cap] */
   uint8_t *level_of_state;	/* [s_cap] */
   uint8_t *domain_type;	/* [s_cap] */
   int16_t *tree;		/* [s_cap][2] */
   uint16_t *x, *y;		/* [s_cap][2] */
   int16_t *into;		/* [s_cap][2][6] */
   float   *weight;		/* [s_cap][2][6] */
   int16_t *y_state;		/* [s_cap][2] */
   uint8_t *y_column;		/* [s_cap][2] */
   /* model snapshots of the DFS */
   int16_t *snap;		/* [FB_MAXDEPTH][2][blob_len] */
   struct TileResult *result;
   fb200_trace_rec_t *trace;	/* [trace_cap] or NULL */
   /* predicted frames only */
   const int16_t *past;		/* [width*height] regenerated reference frame */
   const int16_t *future;	/* the same for the backward prediction of a B frame */
   float   *T2;			/* [tn][s_cap]	 products of the nested (prediction error) pass */
   float   *norms;		/* [1 or 2][p_max - p_min + 1][4 sr^2] norms tables of the motion search
				   (forward; B frames: then backward) */
   float   *pix2;		/* [2^p_max]	 prediction error block, bintree order */
   int	   *norm2;		/* [tn]		 its sums of squares per node */
   uint8_t *saved_dt;		/* [s_cap]	 domain types of the states a prediction attempt hides */
   int8_t  *mv_type, *mv_fx, *mv_fy;	/* [s_cap][2] motion vectors of the ranges (wfa->mv_tree) */
   int8_t  *mv_bx, *mv_by;
   uint8_t *yc_ref;		/* [FB200_MAXSTATES][2] or NULL: the reference's wfa->y_column array as the frames
				   of the sequence so far have left it, by the reference's state numbers.  The
				   reference writes the entries of a colour frame's three virtual states to the
				   stream without ever setting them (output/matrices.c:491): stale values of states
				   that had those numbers before, in this frame or an earlier one */
};

struct TileResult
{
   int	    status;
   unsigned states, basis_states, root_state;
   float    costs [3], err [3], tree_bits [3], matrix_bits [3], weights_bits [3];
   int	    band_root [3];
   int	    trace_len;
   int	    lc_min_end;		/* lc_min_level after the frame (raised by the chroma bands) */
   unsigned progress [3][4];	/* percent values the reference's progress meter shows, per band */
   unsigned long long mp_calls, mp_steps, pass2, blocks, ip_bytes;
   unsigned long long mp_bytes, ss_bytes;	/* algorithmic bytes of the other phases */
   unsigned long long cyc_total, cyc_T, cyc_mp, cyc_append; /* SM cycles per phase */
   unsigned long long lap [16];	/* thread-0 lap timer per sub-phase */
};

size_t fb_tile_kernel_smem (const DevParams &p, int nt);
int    fb_tile_kernel_threads (const DevParams &p, int n_tiles);
int    fb_tile_kernel_cluster (const DevParams &p, int n_tiles);	/* blocks per stream */
int    fb_tile_kernel_occupancy (const DevParams &p);	/* resident tiles per SM */
cudaError_t fb_launch_tile_kernel (const DevParams &p, const TileWs *d_ws, int n_tiles,
				   cudaStream_t stream);
cudaError_t fb_launch_probe (int kind, int n, const float *f, const int *a, const int *b,
			     const int *c, int *out_i, float *out_f);

#endif
