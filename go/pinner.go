//go:build cuda

package main

import (
	"runtime"
	"unsafe"
)

// runtimePinner wraps runtime.Pinner (Go 1.21+): grb_mesh_desc carries pointers to Go slices
// inside a C struct, which cgo only allows when the pointees are pinned for the call.
type runtimePinner struct{ p runtime.Pinner }

func (r *runtimePinner) ptr(p unsafe.Pointer) unsafe.Pointer {
	r.p.Pin(p)
	return p
}

func (r *runtimePinner) Unpin() { r.p.Unpin() }
