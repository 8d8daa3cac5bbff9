// Entry points declared in include/rocco_b200.h whose kernels are not written yet.
// They fail loudly (status -3 + message); nothing here computes anything.
#include "common.cuh"
#define RB_API extern "C" __attribute__((visibility("default")))
#define NOT_YET(name) do { rb::set_error(name ": kernel not implemented yet"); return rb::ST_CUDA; } while (0)

RB_API int rocco_b200_chain_sweep_dev(const double *, size_t, double, const double *, int, long long *, double *, double *, void *) { NOT_YET("rocco_b200_chain_sweep_dev"); }
RB_API int rocco_column_stat_f64(const double *, size_t, size_t, int, double, double, double, double *) { NOT_YET("rocco_column_stat_f64"); }
RB_API int rocco_b200_column_stat_dev(const void *, int, size_t, size_t, int, double, double, double, double *, void *) { NOT_YET("rocco_b200_column_stat_dev"); }
