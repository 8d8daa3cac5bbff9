// Entry points declared in include/rocco_b200.h whose kernels are not written yet.
// They fail loudly (status -3 + message); nothing here computes anything.
#include "common.cuh"
#define RB_API extern "C" __attribute__((visibility("default")))
#define NOT_YET(name) do { rb::set_error(name ": kernel not implemented yet"); return rb::ST_CUDA; } while (0)

RB_API int rocco_b200_chain_sweep_dev(const double *, size_t, double, const double *, int, long long *, double *, double *, void *) { NOT_YET("rocco_b200_chain_sweep_dev"); }
RB_API int rocco_crossfit_whittaker_baseline_f64(const double *, size_t, double, double *) { NOT_YET("rocco_crossfit_whittaker_baseline_f64"); }
RB_API int rocco_crossfit_whittaker_baseline_matrix_f64(const double *, size_t, size_t, double, double *) { NOT_YET("rocco_crossfit_whittaker_baseline_matrix_f64"); }
RB_API int rocco_score_centered_wls_f64(const double *, size_t, size_t, double, double, double, int, int, double, double *, double *, double *, double *, double *, double *, double *, int *) { NOT_YET("rocco_score_centered_wls_f64"); }
RB_API void rocco_b200_default_score_params(rocco_b200_score_params *p)
{
    if (!p) return;
    p->lower_bound_z = 1.0; p->prior_df = 5.0; p->min_effect = 0.0; p->use_min_effect = 0;
    p->spatial_window = 31; p->precision_floor_ratio = 0.01; p->baseline_window = 101; p->reserved = 0;
}
RB_API int rocco_score_loci_wls_f64(const double *, size_t, size_t, const rocco_b200_score_params *, rocco_b200_score_outputs *) { NOT_YET("rocco_score_loci_wls_f64"); }
RB_API int rocco_score_loci_wls_f32(const float *, size_t, size_t, const rocco_b200_score_params *, rocco_b200_score_outputs *) { NOT_YET("rocco_score_loci_wls_f32"); }
RB_API int rocco_b200_score_loci_wls_dev(const void *, int, size_t, size_t, const rocco_b200_score_params *, rocco_b200_score_outputs *, void *) { NOT_YET("rocco_b200_score_loci_wls_dev"); }
RB_API int rocco_b200_crossfit_baseline_dev(const double *, size_t, size_t, double, double *, void *) { NOT_YET("rocco_b200_crossfit_baseline_dev"); }
RB_API int rocco_b200_score_centered_wls_dev(const double *, size_t, size_t, const rocco_b200_score_params *, rocco_b200_score_outputs *, void *) { NOT_YET("rocco_b200_score_centered_wls_dev"); }
RB_API int rocco_column_stat_f64(const double *, size_t, size_t, int, double, double, double, double *) { NOT_YET("rocco_column_stat_f64"); }
RB_API int rocco_b200_column_stat_dev(const void *, int, size_t, size_t, int, double, double, double, double *, void *) { NOT_YET("rocco_b200_column_stat_dev"); }
