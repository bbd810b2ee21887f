// placeholder until corto_decode.cu lands
extern "C" const char *uvol_corto_stage_name(int) { return ""; }
struct CortoBatch {};
void uvol_corto_batch_free(CortoBatch *b) { delete b; }
