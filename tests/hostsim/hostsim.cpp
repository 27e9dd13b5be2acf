// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE: compiles the kernels' __host__ __device__ stage
// bodies with g++ so their logic can be checked against the reference oracle on a CPU-only box.
// Never linked into the product library; the product path only launches the CUDA kernels.
#include <string.h>
#include <stdlib.h>
#include "../../esp32-opus-player_b200/csrc/celt_parse.cuh"

using namespace opusb200;

extern "C" {

int hs_celt_parse(const uint8_t* pkt, int len, int C, int LM, int start, int disable_inv, CeltFrameHdr* hdr, BandOp* ops,
                  int16_t* X) {
    static int16_t leaf[kMaxLeaf];
    ParseCtx p;
    memset(&p, 0, sizeof(p));
    rd_init(p.rd, pkt, (u32)len);
    p.C = C; p.LM = LM; p.start = start; p.end = kNbEBands; p.disable_inv = disable_inv;
    p.ops = ops; p.ops_cap = kMaxOps; p.Xg = X; p.leaf.p = leaf; p.leaf.stride = 1;
    return celt_parse_frame(p, *hdr);
}
int hs_sizeof_hdr() { return (int)sizeof(CeltFrameHdr); }
int hs_sizeof_op() { return (int)sizeof(BandOp); }

}
