// placeholder until corto_decode.cu lands
extern "C" const char *uvol_corto_stage_name(int) { return ""; }
