// Host interface of the persistent single-sequence LSTM layer kernel (rnn_persistent.cu).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace empose {

struct LstmPersistentParams {
    const __half* w_hh[2];     // per direction: [4H][H] fp16, torch gate order i|f|g|o
    const float* xw[2];        // per direction: [F][4H] input projection + both biases (torch gate order)
    const float* h0;           // [dirs][H] initial hidden state or null (zeros)
    const float* c0;           // [dirs][H] or null
    float* h_out;              // [dirs][H] final state or null (may alias h0 / c0)
    float* c_out;
    float* hseq;               // layer output [F][hseq_pitch] elements of `hseq_mode`: columns [dir*H, dir*H + H)
    int64_t hseq_pitch;
    int hseq_mode;             // OperandMode
    float* hx;                 // scratch [dirs][2][H] fp32: the hidden vector handed from step to step through L2
    unsigned* counters;        // scratch [2][32]
    int F, len, H, dirs, C, U; // C CTAs per direction, U = H / C hidden units per CTA
};

// largest C (<= num_sms / dirs, dividing H) whose W_hh slice fits in shared memory; false if none
bool lstm_persistent_pick(int H, int dirs, int num_sms, int* C_out, int* U_out);
size_t lstm_persistent_smem_bytes(int H, int U);
int launch_lstm_persistent(const LstmPersistentParams& p, cudaStream_t s);

}  // namespace empose
