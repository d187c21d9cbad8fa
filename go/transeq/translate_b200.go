//go:build b200

// Package transeq: the B200 build of Translate.
//
// Drop this file and warning_b200.go next to the reference's transeq/*.go, give the stock
// transeq.go the inverse build tag (//go:build !b200) and build with
//
//	CGO_ENABLED=1 go build -tags b200 ./...
//
// Options (options.go), main.go and transeq_test.go stay untouched: TestAllOptions
// (transeq_test.go:14-62) then runs the 29 golden cases through the GPU path.
//
// NOT built in this repository's image (no Go toolchain).  tests/c/stream_drive.c drives the same C
// call sequence on the device (-m gpu), and gotranseq_b200/host/transeq.cpp / gotranseq_b200/transeq.py
// are the compiled / interpreted mirrors of this loop.
package transeq

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gotranseq_b200 -lgotranseq_b200
#include <stdlib.h>
#include "gotranseq_b200.h"

// defined in warning_b200.go (//export): prints the WARNING line of encodedSequence.go:39
extern void gtsGoWarning(void *user, uint8_t *header, size_t header_len, uint8_t b, uint64_t pos);

static int gts_install_go_warning(gts_ctx *ctx) {
    return gts_set_warning_callback(ctx, (gts_warning_fn)gtsGoWarning, NULL);
}
*/
import "C"

import (
	"context"
	"errors"
	"io"
	"unsafe"
)

// AnyOrder lets the stream hand results back as they complete instead of in input order: what the
// reference does with NumWorker > 1, where every worker writes its own buffer when it is full
// (writer.go:171-173).  Whole records, each exactly once.  Off by default: the output is then the
// reference's -n 1 output byte for byte.  (A package variable: Options must stay the reference's struct.)
var AnyOrder = false

// Translate keeps the reference's signature (transeq.go:114).  NumWorker is accepted and ignored:
// every visible GPU is used and the output is in input order (the reference's -n 1 order) unless AnyOrder is set.
func Translate(inputSequence io.Reader, out io.Writer, options Options) error {
	return TranslateContext(context.Background(), inputSequence, out, options)
}

// TranslateContext is Translate with a caller-owned context.  The reference creates its own
// cancellable context (transeq.go:130) and cancels it on the first write error
// (writer.go:171-179); callers had no handle on it.  Here cancelling cctx stops the stream.
func TranslateContext(cctx context.Context, inputSequence io.Reader, out io.Writer, options Options) error {
	frame := C.CString(options.Frame)
	defer C.free(unsafe.Pointer(frame))
	opt := C.gts_options{
		frame: frame, table: C.int32_t(options.Table),
		clean: b2u(options.Clean), alternative: b2u(options.Alternative), trim: b2u(options.Trim),
		num_worker: C.int32_t(options.NumWorker), device: -1,
		num_gpus: -1, // every visible GPU
	}
	if AnyOrder {
		opt.any_order = 1
	}
	var ctx *C.gts_ctx
	if rc := C.gts_create(&opt, &ctx); rc != C.GTS_OK {
		return errors.New(C.GoString(C.gts_last_error(nil))) // same texts as transeq.go:107 / code.go:378
	}
	defer C.gts_destroy(ctx)
	C.gts_install_go_warning(ctx)

	// cancellation from outside: gts_cancel is safe from any thread
	stop := make(chan struct{})
	defer close(stop)
	go func() {
		select {
		case <-cctx.Done():
			C.gts_cancel(ctx)
		case <-stop:
		}
	}()
	lastErr := func() error {
		if err := cctx.Err(); err != nil {
			return err
		}
		return errors.New(C.GoString(C.gts_last_error(ctx)))
	}

	for eof := false; !eof; {
		var buf *C.uint8_t
		var capacity C.size_t
		if rc := C.gts_acquire_input(ctx, &buf, &capacity); rc != C.GTS_OK {
			return lastErr()
		}
		// pinned, C-owned memory: no Go pointer is retained by C
		n, err := io.ReadFull(inputSequence, unsafe.Slice((*byte)(unsafe.Pointer(buf)), int(capacity)))
		eof = err != nil // io.EOF / io.ErrUnexpectedEOF end the stream; other read errors are
		// swallowed exactly like the reference's unchecked scanner.Err() (transeq.go:192-215)
		e := C.int(0)
		if eof {
			e = 1
		}
		if rc := C.gts_submit(ctx, C.size_t(n), e); rc != C.GTS_OK {
			return lastErr()
		}
		for {
			var p *C.uint8_t
			var m C.size_t
			if rc := C.gts_next_output(ctx, &p, &m); rc != C.GTS_OK {
				return lastErr()
			}
			if m == 0 {
				break
			}
			if _, werr := out.Write(unsafe.Slice((*byte)(unsafe.Pointer(p)), int(m))); werr != nil {
				// writer.go:171-179: "fail to write to output file: %v", everything queued is dropped
				msg := C.CString(werr.Error())
				C.gts_report_write_error(ctx, msg)
				C.free(unsafe.Pointer(msg))
				return errors.New(C.GoString(C.gts_last_error(ctx)))
			}
			C.gts_release_output(ctx)
		}
	}
	return nil
}

func b2u(b bool) C.uint8_t {
	if b {
		return 1
	}
	return 0
}
