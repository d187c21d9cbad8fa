//go:build b200

package transeq

/*
#include "gotranseq_b200.h"
*/
import "C"

import (
	"os"
	"unsafe"
)

// gtsGoWarning is the gts_warning_fn of the shim: one call per invalid nucleotide, in input order,
// on the goroutine that drives gts_next_output.  gts_format_warning produces exactly the bytes the
// reference's fmt.Printf writes (encodedSequence.go:39), including the four raw length bytes in
// front of the header.
//
//export gtsGoWarning
func gtsGoWarning(user unsafe.Pointer, header *C.uint8_t, headerLen C.size_t, b C.uint8_t, pos C.uint64_t) {
	need := C.gts_format_warning(header, headerLen, b, pos, nil, 0)
	line := make([]byte, int(need))
	C.gts_format_warning(header, headerLen, b, pos, (*C.char)(unsafe.Pointer(&line[0])), need)
	os.Stdout.Write(line)
}
