// Persistent GRU layer, backward (placeholder until the kernel lands)
#include "gru_persist.cuh"
namespace ipn {
bool gru_persist_bwd_shape_ok(const IpnGruLayerBwd*) { return false; }
long long gru_persist_bwd_ws_bytes(const IpnGruLayerBwd*) { return 0; }
int gru_persist_bwd(const IpnGruLayerBwd*, void*, long long, cudaStream_t) { return IPN_ERR_ARG; }
}
