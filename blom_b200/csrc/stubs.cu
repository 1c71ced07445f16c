// Entry points of include/blomgpu.h whose kernels are not written yet fail
// loudly (there is no CPU fallback).  Each stub disappears when its routine lands.
#include "common.cuh"
namespace blom {
#define NOT_YET(sig, what) void sig { throw std::runtime_error("blomgpu: " what " is not implemented in this build"); }
NOT_YET(eddtra_dev(int, int, int, int, int, int), "eddtra")
NOT_YET(pbcor1_dev(int, int, int, int, int, int), "pbcor1")
NOT_YET(pbcor2_dev(int, int, int, int, int, int), "pbcor2")
}
