// blueice_b200 -- instantiations of the grouped K2 kernel for C = 2 corners, S = 1..8 sources.
#include "bi_unbinned_grouped.cuh"
BI_DEFINE_GROUPED_TU(2)
