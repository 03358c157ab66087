// TEST INFRASTRUCTURE (oracle build only) -- not part of the product.
//
// Lua-free replacement for the reference's Debug.cpp (reference: h264/Debug.cpp:41-122,
// which embeds Lua 5.1 -- not installed in this image).  It defines exactly the four
// out-of-line members declared in h264/Debug.h:67-72 and leaves all 19 print switches
// at zero, i.e. the state the reference's own Debug_conf.lua produces by default.
#include "Debug.h"
#include <cstring>
#include <cstdlib>

Debug::Debug(const char*)
{
    upr_t = cur_t = clock();
    de_cabac_state_running = de_cabac_result_ae = de_cabac_result_bin = 0;
    de_cabac_result_residual = de_residual_transcoeff = de_macroblcok = 0;
    de_residual_result = de_residual_result_Y = de_residual_result_Cb = de_residual_result_Cr = 0;
    de_prediction_result = de_prediction_result_Y = de_prediction_result_Cb = de_prediction_result_Cr = 0;
    de_conspic_result = de_conspic_result_Y = de_conspic_result_Cb = de_conspic_result_Cr = 0;
    de_inter_movevector = de_pic_terminalchar = de_timer = de_nal_info = 0;
    control_all = true;
    // H264O_TRACE_AE=1: print every syntax-element value the arithmetic decoder returns (cabac::cread_ae, h264/cabac.cpp:43-50) --
    // how a divergence between a written stream and the reference's reading of it is located
    if (getenv("H264O_TRACE_AE")) de_cabac_result_ae = 1;
}
Debug::~Debug() {}
double Debug::get_RunTime() { return (double)clock() / CLOCKS_PER_SEC; }
double Debug::de_DltTime(const char*)
{
    cur_t = clock();
    double d = (double)(cur_t - upr_t) / CLOCKS_PER_SEC;
    upr_t = cur_t;
    return d;
}
