// idw_i8.cuh -- interface of the tcgen05 (int8-sliced fp64) IDW kernel, see idw_i8.cu.
#pragma once
#include "common.cuh"

namespace smrfb {

constexpr int IDW_I8_MAX_STATIONS = 224;   // 7 K-blocks of 32 stations
constexpr int IDW_I8_STEPS = 48;           // timesteps per chunk (the MMA N dimension)

struct IdwI8Scratch {
    uint8_t* d_blob = nullptr;   // sliced residual / validity operands, one blob per chunk of steps
    size_t blob_cap = 0;
};

// One-off per process and device: opt in to the kernel's dynamic shared memory.
int idw_i8_prepare(int device);

// Distribute T steps over the cell window [cbase, cbase + ncw).  rec / valid are prep_steps_kernel's outputs
// (rec[t][RS]: residuals, p0, p1 at S_pad; valid[t][S]).
int idw_i8_run(cudaStream_t st, IdwI8Scratch* scratch, const Geom& g, const double* d_mx, const double* d_my, int S,
               double power, const double* d_rec, int RS, int S_pad, const uint8_t* d_valid, int T, int detrend,
               double vmin, double vmax, double* d_out, int out_f32, long long cbase, long long ncw, int device);

}  // namespace smrfb
