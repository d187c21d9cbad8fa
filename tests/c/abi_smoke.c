/* The header must be usable from plain C (this is what a cgo shim compiles against), and the
 * option-decoding entry points must work without a device.  Built and run by tests/test_abi.py. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "gotranseq_b200.h"

int main(void) {
    uint32_t mask = 0;
    int reverse = -1;
    char aa[65];
    uint16_t lut[125];
    gts_options opt;
    gts_ctx *ctx = NULL;
    gts_piece piece;
    gts_piece_info info;
    gts_segment segs[GTS_MAX_SEGMENTS];

    memset(&opt, 0, sizeof(opt));
    memset(&piece, 0, sizeof(piece));
    memset(&info, 0, sizeof(info));
    memset(segs, 0, sizeof(segs));
    if (gts_frame_mask("6", &mask, &reverse) != GTS_OK || mask != 63 || reverse != 1) return 1;
    if (gts_frame_mask("7", &mask, &reverse) != GTS_ERR_FRAME) return 2;
    memset(aa, 0, sizeof(aa));
    if (gts_codon_table(11, aa) != GTS_OK || strlen(aa) != 64) return 3;
    if (gts_codon_table(1, aa) != GTS_ERR_TABLE) return 4;
    if (gts_fused_lut(0, 1, lut) != GTS_OK) return 5;
    /* an unknown frame is rejected before any device is touched, with the reference's message */
    opt.frame = "9";
    opt.device = -1;
    if (gts_create(&opt, &ctx) != GTS_ERR_FRAME || ctx != NULL) return 6;
    if (strstr(gts_last_error(NULL), "wrong value for -f | --frame parameter: 9") == NULL) return 7;
    opt.any_order = 1; /* (the reference's -n > 1 order; a plain int32 at the end of the options) */
    printf("abi ok %s %u bytes/piece_info %u bytes/piece options %u any_order %u\n", GTS_VERSION, (unsigned)sizeof(info),
           (unsigned)sizeof(piece), (unsigned)sizeof(opt), (unsigned)offsetof(gts_options, any_order));
    return 0;
}
