/* placeholder until the UASTC restatement lands (no fixture, no encoder) */
#include <stdint.h>
int uvo_uastc_block_to_rgba(const uint8_t *blk, uint8_t *rgba) { (void)blk; (void)rgba; return -3; }
