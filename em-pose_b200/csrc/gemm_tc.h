// Host interface of the tcgen05 job executor (gemm_tc.cu) and the FFMA job executor (gemm_simt.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_jobs.h"

namespace empose {

constexpr int kTensorMapBytes = 128;   // sizeof(CUtensorMap)

// Encode a 2-D tensor map {k_extent, rows} of fp32 (half = 0) or fp16 (half = 1) elements with row stride
// `row_stride_elems`, box {128 bytes, box_rows}, 128-byte swizzle, zero fill out of bounds.  `out_map` is HOST
// memory of kTensorMapBytes.
int tc_encode_map(void* out_map, const float* base, int64_t row_stride_elems, int k_extent, int64_t rows,
                  int box_rows, int half);

// Run jobs [job_begin, job_begin + job_count) of the device array `d_jobs` on `m_tiles` row tiles.
// `jobs_per_item` consecutive jobs form one work item executed by one CTA in order.
int tc_launch(const GemmJob* d_jobs, const void* d_maps, int job_begin, int job_count, int jobs_per_item, int m_tiles,
              int num_sms, cudaStream_t stream);

// The persistent wavefront form: ONE launch executes `n_items` (job, row-tile unit) pairs given by the device table
// `d_items` (int2 {job index, unit}) in table order, CTAs taking items round-robin; jobs order themselves across CTAs
// through GemmJob::wait_ctr / done_ctr, so every item may only wait for items EARLIER in the table.  A unit is
// tc_item_rows() rows: one 128-row tile, or two when the executor runs CTA pairs.  `epoch` = 1, 2, 3, ... counts the
// launches that have used the same counters (they are never reset).
int tc_item_rows(int m_tiles, int num_sms);
int tc_launch_items(const GemmJob* d_jobs, const void* d_maps, const void* d_items, int n_items, int rows_per_unit, uint32_t epoch,
                    int m_tiles, int num_sms, cudaStream_t stream);

// Same contract on the fp32 FFMA executor (one launch per job; tensor maps unused).
int simt_launch(const GemmJob* d_jobs, const GemmJob* h_jobs, int job_begin, int job_count, int m_tiles,
                cudaStream_t stream, int64_t* launch_counter);

}  // namespace empose
