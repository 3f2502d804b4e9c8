// Instantiation of the aggregation kernels for T = __nv_bfloat16, 1 element(s) per lane access.
#include "pna_aggregate_impl.cuh"
namespace pna {
template int launch_typed<__nv_bfloat16, 1>(const KParams&, cudaStream_t);
}
