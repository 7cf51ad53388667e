// cssm_series.cu -- the single-launch series kernels of cssm_series.cuh as a translation unit of their own.
//
// ptxas chooses different code for the same kernel depending on what else its module contains:
// compiled together with the step kernels, either the step kernels or the series kernel came out in their slower
// variant.  Separate modules make the two choices independent; nothing is shared across the boundary but the two
// look-up functions below (cooperative launches take the kernels by host address).
#include "cssm_series.cuh"

namespace cssm {

template <typename real, int ITEMS>
static void* small_ptr(int d, int kind) {
  const bool strat = kind == CSSM_RESAMPLE_STRATIFIED;
  if (d == 2) return strat ? (void*)k_series_small<real, 2, CSSM_RESAMPLE_STRATIFIED, ITEMS> : (void*)k_series_small<real, 2, CSSM_RESAMPLE_SYSTEMATIC, ITEMS>;
  return strat ? (void*)k_series_small<real, 0, CSSM_RESAMPLE_STRATIFIED, ITEMS> : (void*)k_series_small<real, 0, CSSM_RESAMPLE_SYSTEMATIC, ITEMS>;
}
// items = particles per thread: 1 (256-particle tiles) or 2 (512-particle tiles)
void* series_small_kernel(int dtype, int items, int d, int resample_kind) {
  // items == 1 (256-particle tiles, two blocks per SM) measured slower than items == 2 (8.70 vs 8.45 us per observation
  // at 2^16 particles: twice the arrivals per barrier, the same chain of dependent instructions per warp): not instantiated
  (void)items;
  return dtype == CSSM_F32 ? small_ptr<float, 2>(d, resample_kind) : small_ptr<double, 2>(d, resample_kind);
}

// tiny clouds: one block
template <typename real>
static void* one_ptr(int d, int kind) {
  const bool strat = kind == CSSM_RESAMPLE_STRATIFIED;
  if (d == 1) return strat ? (void*)k_series_one<real, 1, CSSM_RESAMPLE_STRATIFIED> : (void*)k_series_one<real, 1, CSSM_RESAMPLE_SYSTEMATIC>;
  if (d == 2) return strat ? (void*)k_series_one<real, 2, CSSM_RESAMPLE_STRATIFIED> : (void*)k_series_one<real, 2, CSSM_RESAMPLE_SYSTEMATIC>;
  return strat ? (void*)k_series_one<real, 0, CSSM_RESAMPLE_STRATIFIED> : (void*)k_series_one<real, 0, CSSM_RESAMPLE_SYSTEMATIC>;
}
void* series_one_kernel(int dtype, int d, int resample_kind) {
  return dtype == CSSM_F32 ? one_ptr<float>(d, resample_kind) : one_ptr<double>(d, resample_kind);
}

// mid-size clouds: several tiles per block
template <typename real, int ITEMS>
static void* multi_ptr(int d, int kind) {
  const bool strat = kind == CSSM_RESAMPLE_STRATIFIED;
  if (d == 7) return strat ? (void*)k_series_multi<real, 7, CSSM_RESAMPLE_STRATIFIED, ITEMS> : (void*)k_series_multi<real, 7, CSSM_RESAMPLE_SYSTEMATIC, ITEMS>;
  return strat ? (void*)k_series_multi<real, 0, CSSM_RESAMPLE_STRATIFIED, ITEMS> : (void*)k_series_multi<real, 0, CSSM_RESAMPLE_SYSTEMATIC, ITEMS>;
}
void* series_multi_kernel(int dtype, int items, int d, int resample_kind) {
  if (dtype == CSSM_F32) return items == 8 ? multi_ptr<float, 8>(d, resample_kind) : multi_ptr<float, 2>(d, resample_kind);
  return items == 8 ? multi_ptr<double, 8>(d, resample_kind) : multi_ptr<double, 2>(d, resample_kind);
}

}  // namespace cssm
