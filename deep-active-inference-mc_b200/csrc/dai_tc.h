// Tensor-core (tcgen05 / TMEM / TMA) path of the contraction layers — internal interface.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "dai_kernels.h"

namespace dai {

struct TcWeights {
    void* impl = nullptr;   // opaque (dai_tc.cu)
};

// Plans the bf16 hi/lo K-major operand images of the contraction layers: appends one RepackJob per packed image
// (dai_kernels.h) whose `dst` is the pointer inside the opaque TcImpl that the image's device address must be stored in.
// Nothing is packed on the host; the handle runs the jobs on the device whenever their source tensor changes.
int tc_plan_weights(TcWeights* out, std::vector<RepackJob>* jobs, std::string* err);
void tc_set_w4(TcWeights* w, const float* w19_host);     // po_net.19.weight, 288 floats
void tc_set_conv1(TcWeights* w, const float* w_host /*qs_net.0.weight (32,1,3,3)*/, const float* bias_host /*[32]*/);
void tc_release(TcWeights* w);

// FC4 -> ct1 -> ct2 -> ct3 -> pixel terms for one chunk of decoder rows on the tensor cores.
// Returns the number of kernels launched, or -1 (+ *err).
// optional per-kernel event recorder (dai_profile_begin / dai_profile_end)
struct LayerTimer {
    struct Rec { int layer; int rows; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    bool on = false;
    void begin(int layer, int rows, cudaStream_t st) {
        if (!on) return;
        Rec r{layer, rows, nullptr, nullptr};
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, st);
        recs.push_back(r);
    }
    void end(cudaStream_t st) { if (on) cudaEventRecord(recs.back().b, st); }
};

// h3b = FC3 output in K-blocked bf16 hi/lo [plane][kc 32][rows_pad][8]; the chunk starts at row0.
int tc_decoder_chunk(const TcWeights& tw, const DevWeights& w, int precision, const void* h3b, size_t rows_pad, int row0,
                     const uint32_t* mask, int nrows, void* act0, void* act1, void* act2, void* act3, const Ct4Args& c4,
                     cudaStream_t st, std::string* err, LayerTimer* timer = nullptr);

// Hidden dense layers served by the tensor-core dense kernel
enum { TC_PS1 = 0, TC_PS2, TC_PO1, TC_PO2, TC_QS0, TC_QS1, TC_QS2, TC_QC4 };
// bias + ReLU + keyed dropout (site of the row's set + layer); in/out are K-blocked bf16 hi/lo planes
// [plane][K/8][rows_pad][8].  Returns launches or -1.
int tc_dense_hidden(const TcWeights& tw, int which, int precision, const void* in, void* out, int rows, size_t rows_pad,
                    const NoiseKey& nk, const NoiseRows& nr, int layer, cudaStream_t st, std::string* err);

// Encoder conv2 (32->32, 31x31->15x15) and conv3 (32->64, 15x15->7x7) on tensor cores.  c1 / c2 are parity-split
// channel-blocked bf16 hi/lo planes, c3 is fp32 NHWC (rows,7,7,64).  Returns launches or -1.
// img != null: conv1 (1->32, 64x64 -> 31x31) is computed inside conv2's kernel from the fp32 images (rows,64,64) and c1 is
// not read (it must still be a valid buffer of the c1 size); img == null: c1 holds conv1's output (launch_qs_conv1).
int tc_qs_convs(const TcWeights& tw, const DevWeights& w, int precision, const void* c1, void* c2, float* c3, int rows,
                cudaStream_t st, std::string* err, const float* img = nullptr);

// Encoder conv4 (64->64, k3 s2, 7x7 -> 3x3) as im2col + tcgen05 GEMM.  c3 fp32 NHWC (rows,7,7,64); out = the K-blocked
// bf16 hi/lo operand of the encoder's FC1 ([plane][72][rows_pad][8], k = pixel*64 + c).  Returns launches or -1.
size_t tc_qs_conv4_scratch_bytes(int rows);
int tc_qs_conv4(const TcWeights& tw, int precision, const float* c3, int rows, void* scratch, size_t rows_pad, void* out,
                cudaStream_t st, std::string* err);

// ct2 -> ct3 fused on CTA pairs (k_tc_ct23): act1 blocked planes in, the last deconv's row planes out; `scratch` =
// tc_ct23_scratch_bytes(nrows) bytes that stay L2-resident (2 images per CTA).
size_t tc_ct23_scratch_bytes(int nrows);
int tc_ct23(const TcWeights& tw, const DevWeights& w, int precision, const void* act1, void* scratch, void* act3, int nrows,
            cudaStream_t st, std::string* err);

// One tensor-core layer (1: ct1, 2: ct2, 3: ct3) on channel-blocked bf16 hi/lo input planes.
int tc_layer(const TcWeights& tw, const DevWeights& w, int precision, int layer, const void* in, void* out, int nrows,
             cudaStream_t st, std::string* err);

// fp32 NHWC [rows][HW][C] <-> channel-blocked bf16 planes [hi|lo][rows][C/8][HW][8] (test / debug converters)
int tc_to_blocked(const float* nhwc, int rows, int hw, int C, void* blocked, cudaStream_t st);
int tc_from_blocked(const void* blocked, int rows, int hw, int C, float* nhwc, cudaStream_t st);

}  // namespace dai
