// grid.cu — node store + uniform-grid spatial hash + exact k-NN (K2/K3).  (under construction)
#include "common.cuh"
void rtpp_nodes_destroy(rtpp_ctx *ctx) { (void)ctx; }
