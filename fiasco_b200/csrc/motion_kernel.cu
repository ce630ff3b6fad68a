/*
 *  motion_kernel.cu -- the norms tables of the motion search for sm_100a.
 *
 *  For a predicted frame the reference computes, lazily and block by block while it walks
 *  the bintree, the squared norm of the motion compensated prediction error of every block
 *  of level p_min_level for every displacement of the search window (fill_norms_table,
 *  codec/mwfa.c:544-602, mcpe_norm :651-684): 4 * search_range^2 (= 1024) sums of
 *  ((original - displaced reference) / 16)^2 per block.  Nothing in those numbers depends on
 *  the coder's state, so the B200 form computes them for ALL blocks of the frame in one
 *  launch before the recursion starts -- the one part of the codec that is shaped by memory
 *  bandwidth rather than by a dependent chain: per block the original pixels and the
 *  (block + 2 * search_range)^2 reference window are staged in shared memory once and
 *  every displacement reads its sub-window from there; one thread sums one displacement in
 *  the reference's row order in fp32 (integer terms, int16 difference with wrap-around,
 *  truncating division), so the tables are bit-identical whatever the level.
 *
 *  Output layout: norms [by][bx][index], index = (my + sr) * 2 sr + (mx + sr) as in the
 *  reference's loops; blocks that are not completely inside the frame, and displacements
 *  that leave the frame, hold 0 (mwfa.c:569-576).
 */
#include <stdint.h>
#include <cuda_runtime.h>
#include "fiasco_b200.h"
#include "tile_kernel.cuh"

namespace {

__global__ void __launch_bounds__ (256)
fiasco_norms_kernel (const int16_t *orig, const int16_t *past, int width, int height,
		     int bw, int bh, int sr, float *norms)
{
   FB_DYN_SMEM (int16_t, sm);
   const int bx = blockIdx.x, by = blockIdx.y;
   const int x0 = bx * bw, y0 = by * bh;
   const int ww = bw + 2 * sr, wh = bh + 2 * sr;	/* reference window */
   int16_t  *so = sm;					/* [bh][bw] original block */
   int16_t  *sw = sm + bw * bh;				/* [wh][ww] reference window */
   const int nd = 4 * sr * sr;
   float    *out = norms + ((size_t) by * gridDim.x + bx) * nd;

   if (x0 + bw > width || y0 + bh > height)	/* uniform: the block is not a candidate */
   {
      for (int i = threadIdx.x; i < nd; i += blockDim.x)
	 out [i] = 0.0f;
      return;
   }
   for (int i = threadIdx.x; i < bw * bh; i += blockDim.x)
      so [i] = orig [(size_t) (y0 + i / bw) * width + x0 + i % bw];
   for (int i = threadIdx.x; i < ww * wh; i += blockDim.x)
   {
      const int x = x0 - sr + i % ww, y = y0 - sr + i / ww;

      sw [i] = (x >= 0 && x < width && y >= 0 && y < height) ? past [(size_t) y * width + x] : (int16_t) 0;
   }
   __syncthreads ();
   for (int index = threadIdx.x; index < nd; index += blockDim.x)
   {
      const int mx = index % (2 * sr) - sr, my = index / (2 * sr) - sr;
      float	norm = 0.0f;

      if (x0 + mx >= 0 && x0 + mx + bw <= width && y0 + my >= 0 && y0 + my + bh <= height)
	 for (int y = 0; y < bh; y++)
	    for (int x = 0; x < bw; x++)
	    {
	       const int16_t d = (int16_t) (so [y * bw + x] - sw [(y + my + sr) * ww + x + mx + sr]);
	       const int     q = d / 16;

	       norm += (float) (q * q);
	    }
      out [index] = norm;
   }
}

} /* namespace */

#define MN_TRY(call)                                                                      \
   do {                                                                                   \
      cudaError_t e_ = (call);                                                            \
      if (e_ != cudaSuccess) {                                                            \
	 if (err && errlen)                                                               \
	    snprintf (err, errlen, "CUDA error %s (%s)", cudaGetErrorName (e_),           \
		      cudaGetErrorString (e_));                                           \
	 cudaFree (d_orig); cudaFree (d_past); cudaFree (d_norms);                        \
	 return FB200_ECUDA;                                                              \
      }                                                                                   \
   } while (0)

#include <stdio.h>

extern "C" int
fb200_motion_norms (int device, const int16_t *orig, const int16_t *past, int width,
		    int height, int level, int search_range, float *norms, float *kernel_ms,
		    char *err, size_t errlen)
{
   int16_t *d_orig = NULL, *d_past = NULL;
   float   *d_norms = NULL;

   if (!orig || !past || !norms || width < 2 || height < 2 || level < 2 || level > 12
       || search_range < 1 || search_range > 16)
   {
      if (err && errlen)
	 snprintf (err, errlen, "fb200_motion_norms: bad arguments");
      return FB200_EINVAL;
   }
   if (fb200_device_count () <= device)
   {
      if (err && errlen)
	 snprintf (err, errlen, "no CUDA device %d available (this library has no CPU path)", device);
      return FB200_ENODEVICE;
   }
   const int	bw = 1 << (level >> 1), bh = 1 << ((level + 1) >> 1);
   const int	nbx = (width + bw - 1) / bw, nby = (height + bh - 1) / bh;
   const int	nd  = 4 * search_range * search_range;
   const size_t npix = (size_t) width * height;
   const size_t smem = ((size_t) bw * bh + (size_t) (bw + 2 * search_range) * (bh + 2 * search_range)) * 2;
   cudaEvent_t	e0, e1;

   MN_TRY (cudaSetDevice (device));
   MN_TRY (cudaMalloc (&d_orig, npix * 2));
   MN_TRY (cudaMalloc (&d_past, npix * 2));
   MN_TRY (cudaMalloc (&d_norms, (size_t) nbx * nby * nd * 4));
   MN_TRY (cudaMemcpy (d_orig, orig, npix * 2, cudaMemcpyHostToDevice));
   MN_TRY (cudaMemcpy (d_past, past, npix * 2, cudaMemcpyHostToDevice));
   MN_TRY (cudaFuncSetAttribute (fiasco_norms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
   MN_TRY (cudaEventCreate (&e0));
   MN_TRY (cudaEventCreate (&e1));
   MN_TRY (cudaEventRecord (e0));
   FB_LAUNCH (fiasco_norms_kernel, dim3 (nbx, nby), 256, smem, 0, d_orig, d_past, width, height, bw, bh,
	      search_range, d_norms);
   MN_TRY (cudaGetLastError ());
   MN_TRY (cudaEventRecord (e1));
   MN_TRY (cudaEventSynchronize (e1));
   if (kernel_ms)
      cudaEventElapsedTime (kernel_ms, e0, e1);
   cudaEventDestroy (e0);
   cudaEventDestroy (e1);
   MN_TRY (cudaMemcpy (norms, d_norms, (size_t) nbx * nby * nd * 4, cudaMemcpyDeviceToHost));
   cudaFree (d_orig);
   cudaFree (d_past);
   cudaFree (d_norms);
   return FB200_OK;
}
