// TEST INFRASTRUCTURE (oracle build only) -- not part of the product.
//
// Queue-fed stand-in for the reference's arithmetic decoder (reference: h264/cabac.cpp).
// It implements only the public surface other translation units link against
// (h264/cabac.h:67-74: cread_ae, slice_end, set_pic, set_slice, ctor, dtor).  Instead of
// decoding bins, cread_ae() pops the next pre-recorded syntax-element value from a queue
// that the harness loaded from a side file.  Everything downstream of entropy decoding --
// macroblock::Parse / Init0 / Calc, residual::Parse / Decode, picture, Decoder -- is the
// unmodified reference and runs on exactly the values a CABAC stream would have produced.
//
// Each queue entry carries the syntax code the producer expected the reference to ask
// for; a mismatch means the producer's serialisation order diverged from
// macroblock::Parse (h264/macroblock.cpp:143-475) / residual_block_cabac
// (h264/residual.cpp:561-614) and aborts loudly.
#include "cabac.h"
#include "Parser.h"
#include <cstdio>
#include <cstdlib>
#include <vector>

std::vector<int> g_h264o_queue;      // pairs (code, value)
size_t g_h264o_queue_pos = 0;

static inline int code_class(int syntax)
{
    return syntax >= 0x1000 ? ((syntax >> 12) & 0xFF) : syntax;
}

int cabac::cread_ae(int syntax)
{
    if (g_h264o_queue_pos + 2 > g_h264o_queue.size()) {
        fprintf(stderr, "cabac_queue: queue exhausted at request 0x%x\n", syntax);
        exit(3);
    }
    int code = g_h264o_queue[g_h264o_queue_pos];
    int value = g_h264o_queue[g_h264o_queue_pos + 1];
    if (code != code_class(syntax)) {
        fprintf(stderr, "cabac_queue: entry %zu holds code 0x%x but the reference asked for 0x%x (class 0x%x)\n",
                g_h264o_queue_pos / 2, code, syntax, code_class(syntax));
        exit(4);
    }
    g_h264o_queue_pos += 2;
    return value;
}
bool cabac::slice_end() { state = 0; return true; }
bool cabac::set_pic(picture* pic1) { this->pic = pic1; return true; }
bool cabac::set_slice(Slice* sl) { this->lifeTimeSlice = sl; return true; }
cabac::cabac(Parser* parser)
{
    b3 = 0; b1 = 0; state = 0;
    ctxIdxOfInitVariables = NULL;
    this->p = parser;
    this->debug = parser->debug;
}
cabac::~cabac() {}
