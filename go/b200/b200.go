// Package b200 binds libgoleft_b200.so (CUDA, sm_100a) for goleft's windowed-depth path.
//
// SOURCE ONLY: the build image of this repository has no Go toolchain (`go: command not found`), so this file has never
// been compiled there; the same C ABI (include/goleft_b200.h) is exercised by the C++ CLI (cli/goleft.cpp) and by the
// ctypes binding the tests use (goleft_b200/capi.py).  Build with CGO_ENABLED=1 after `make lib`.
//
// What the reference would call (paths in brentp/goleft v0.2.6):
//   depth/depth.go:392-394  process.Runner(genCommands(args), ...) + callback :238-364
//        -> Bam.DepthContig per reference sequence (decode + count + window means + callable runs + BED text on the GPU)
//   indexcov/indexcov.go:417-434,651-676  Index.init / NormalizedDepth / CountsAtDepth
//        -> Ctx.IndexcovCohort (+ gl_indexcov_counts_batch, gl_indexcov_xnorm, gl_format_g3)
package b200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../goleft_b200 -lgoleft_b200 -Wl,-rpath,${SRCDIR}/../../goleft_b200
#include <stdlib.h>
#include "goleft_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Ctx is one GPU context; not safe for concurrent use (one goroutine at a time, like gl_ctx).
type Ctx struct{ h *C.gl_ctx }

// DeviceCount is the number of visible CUDA devices (goleft depth deals contigs to all of them).
func DeviceCount() int {
	var n C.int
	if C.gl_device_count(&n) != C.GL_OK {
		return 0
	}
	return int(n)
}

// New creates a context on one GPU.  There is no CPU fallback: without a device this fails.
func New(device int) (*Ctx, error) {
	var h *C.gl_ctx
	if rc := C.gl_ctx_create(C.int(device), &h); rc != C.GL_OK {
		return nil, fmt.Errorf("goleft_b200: %s", C.GoString(C.gl_last_error(nil)))
	}
	c := &Ctx{h}
	runtime.SetFinalizer(c, func(c *Ctx) { C.gl_ctx_destroy(c.h) })
	return c, nil
}

func (c *Ctx) err(rc C.int) error {
	if rc == C.GL_OK {
		return nil
	}
	return fmt.Errorf("goleft_b200: %s", C.GoString(C.gl_last_error(c.h)))
}

// LptAssign deals work items (contigs, weight = length) to `bins` GPUs, longest first onto the least loaded one.
func LptAssign(weights []int64, bins int) []int32 {
	out := make([]int32, len(weights))
	if len(weights) == 0 {
		return out
	}
	C.gl_lpt_assign((*C.int64_t)(unsafe.Pointer(&weights[0])), C.int32_t(len(weights)), C.int32_t(bins),
		(*C.int32_t)(unsafe.Pointer(&out[0])), nil)
	return out
}

// Bam is an indexed BAM opened by the library's feeder (BGZF inflate + record parse on the host pool).
type Bam struct{ h *C.gl_bam }

// OpenBam opens path and its .bai.  The index is what `samtools depth -r` needs too (depth/depth.go:116,152).
func OpenBam(path string) (*Bam, error) {
	cs := C.CString(path)
	defer C.free(unsafe.Pointer(cs))
	var h *C.gl_bam
	buf := make([]byte, 512)
	if rc := C.gl_bam_open(cs, &h, (*C.char)(unsafe.Pointer(&buf[0])), 512); rc != C.GL_OK {
		return nil, fmt.Errorf("goleft_b200: %s", C.GoString((*C.char)(unsafe.Pointer(&buf[0]))))
	}
	b := &Bam{h}
	runtime.SetFinalizer(b, func(b *Bam) { C.gl_bam_close(b.h) })
	return b, nil
}

// DepthContig replaces, for one reference sequence, every `samtools depth -Q q -r chrom:b-e` child of
// depth/depth.go:129-159 and the callback of :238-364 run on each of its chunks: it returns the bytes the reference
// appends to <prefix>.depth.bed and <prefix>.callable.bed for that sequence (chunks of `step` bases, depth.go:132).
func (c *Ctx) DepthContig(b *Bam, tid int, chrom string, length int, window, minCov, maxMeanDepth, q int) (depthBed, callableBed []byte, err error) {
	step := 10000000 / window * window
	if step < window {
		step = window
	}
	cs := C.CString(chrom)
	defer C.free(unsafe.Pointer(cs))
	nWin := (length-1)/window + 1
	depthBed = make([]byte, int(C.gl_depth_text_bound(cs, C.int64_t(nWin))))
	callableBed = make([]byte, 1<<20)
	var seg C.gl_bam_segments
	// The feeder on the GPU first (BGZF inflate + record parse + filter + CIGAR walk on the device: only compressed bytes
	// cross PCIe); GL_ESTATE means this reference cannot take that road (no usable index, a member the device inflater
	// rejects) and the host feeder below decodes it.
	var dS, dE *C.int32_t
	var dN C.int64_t
	if rc := C.gl_bam_decode_device(c.h, b.h, C.int32_t(tid), C.int32_t(q), &dS, &dE, &dN, &seg); rc == C.GL_OK {
		if e := c.err(C.gl_depth_begin(c.h, 0, C.int64_t(length))); e != nil {
			return nil, nil, e
		}
		if dN > 0 {
			if e := c.err(C.gl_depth_add_segments_device(c.h, dS, dE, dN)); e != nil {
				return nil, nil, e
			}
		}
		if e := c.err(C.gl_depth_reduce(c.h, C.int32_t(window), C.int32_t(minCov), C.int32_t(maxMeanDepth), C.int64_t(step))); e != nil {
			return nil, nil, e
		}
		for attempt := 0; attempt < 2; attempt++ {
			var dl, cl C.int64_t
			rc := C.gl_depth_text(c.h, cs, (*C.char)(unsafe.Pointer(&depthBed[0])), C.int64_t(len(depthBed)), &dl,
				(*C.char)(unsafe.Pointer(&callableBed[0])), C.int64_t(len(callableBed)), &cl)
			if rc == C.GL_ERANGE && attempt == 0 {
				depthBed = make([]byte, int(dl)+16)
				callableBed = make([]byte, int(cl)+16)
				continue
			}
			if e := c.err(rc); e != nil {
				return nil, nil, e
			}
			return depthBed[:dl], callableBed[:cl], nil
		}
	} else if rc != C.GL_ESTATE {
		return nil, nil, c.err(rc)
	}
	ebuf := make([]byte, 512)
	if rc := C.gl_bam_decode(b.h, C.int32_t(tid), 0, C.int64_t(length), C.int32_t(q), 0, 0, &seg, (*C.char)(unsafe.Pointer(&ebuf[0])), 512); rc != C.GL_OK {
		return nil, nil, fmt.Errorf("goleft_b200: %s", C.GoString((*C.char)(unsafe.Pointer(&ebuf[0]))))
	}
	for attempt := 0; attempt < 2; attempt++ {
		var dl, cl C.int64_t
		var rc C.int
		dp := (*C.char)(unsafe.Pointer(&depthBed[0]))
		cp := (*C.char)(unsafe.Pointer(&callableBed[0]))
		if seg.format == 8 {
			rc = C.gl_depth_bed_contig_packed8(c.h, cs, C.int64_t(length), seg.a0, (*C.uint8_t)(seg.a1), (*C.uint8_t)(seg.a2), seg.n,
				C.int32_t(window), C.int32_t(minCov), C.int32_t(maxMeanDepth), C.int64_t(step),
				dp, C.int64_t(len(depthBed)), &dl, cp, C.int64_t(len(callableBed)), &cl)
		} else {
			rc = C.gl_depth_bed_contig(c.h, cs, C.int64_t(length), seg.a0, (*C.int32_t)(seg.a1), seg.n,
				C.int32_t(window), C.int32_t(minCov), C.int32_t(maxMeanDepth), C.int64_t(step), 0,
				dp, C.int64_t(len(depthBed)), &dl, cp, C.int64_t(len(callableBed)), &cl)
		}
		if rc == C.GL_ERANGE && attempt == 0 {
			depthBed = make([]byte, int(dl)+16)
			callableBed = make([]byte, int(cl)+16)
			continue
		}
		if e := c.err(rc); e != nil {
			return nil, nil, e
		}
		return depthBed[:dl], callableBed[:cl], nil
	}
	return nil, nil, fmt.Errorf("goleft_b200: text buffers did not converge")
}

// DepthRegion replaces one `samtools depth -r chrom:rs+1-re` child plus the per-line loop of the callback
// (depth/depth.go:45,238-364): window sums and class runs for [rs,re).  start/end are the M/=/X blocks of
// the records that pass the flag/MAPQ filter (what biogo's bam.Reader + Cigar walk yields).
func (c *Ctx) DepthRegion(rs, re int, start, end []int32, window, minCov, maxMeanDepth, runBreak int) (sums []int64, runStart []int32, runClass []uint8, err error) {
	nWin := (re-1)/window - rs/window + 1
	sums = make([]int64, nWin)
	cap := 1 << 16
	for {
		runStart = make([]int32, cap)
		runClass = make([]uint8, cap)
		var nw, nr C.int64_t
		var ps, pe *C.int32_t
		if len(start) > 0 {
			ps = (*C.int32_t)(unsafe.Pointer(&start[0]))
			pe = (*C.int32_t)(unsafe.Pointer(&end[0]))
		}
		rc := C.gl_depth_region(c.h, C.int64_t(rs), C.int64_t(re), ps, pe, C.int64_t(len(start)),
			C.int32_t(window), C.int32_t(minCov), C.int32_t(maxMeanDepth), C.int64_t(runBreak),
			(*C.int64_t)(unsafe.Pointer(&sums[0])), C.int64_t(nWin), &nw,
			(*C.int32_t)(unsafe.Pointer(&runStart[0])), (*C.uint8_t)(unsafe.Pointer(&runClass[0])), C.int64_t(cap), &nr)
		if rc == C.GL_ERANGE && int(nr) > cap {
			cap = int(nr) + 1024
			continue
		}
		if e := c.err(rc); e != nil {
			return nil, nil, nil, e
		}
		return sums[:nw], runStart[:nr], runClass[:nr], nil
	}
}

// FormatChunk returns exactly the rows the reference's callback writes for one chunk (depth.go:293-358), BED-mode
// chunk-edge quirks included (host-side; DepthContig formats on the GPU).
func FormatChunk(chrom string, rs, re, window int, sums []int64, runStart []int32, runClass []uint8) (depthBed, callableBed string) {
	cs := C.CString(chrom)
	defer C.free(unsafe.Pointer(cs))
	var d, ca *C.char
	var dl, cl C.int64_t
	C.gl_depth_format_chunk(cs, C.int64_t(rs), C.int64_t(re), C.int32_t(window),
		(*C.int64_t)(unsafe.Pointer(&sums[0])), C.int64_t(len(sums)),
		(*C.int32_t)(unsafe.Pointer(&runStart[0])), (*C.uint8_t)(unsafe.Pointer(&runClass[0])), C.int64_t(len(runStart)),
		&d, &dl, &ca, &cl)
	defer C.gl_free_text(d)
	defer C.gl_free_text(ca)
	return C.GoStringN(d, C.int(dl)), C.GoStringN(ca, C.int(cl))
}

// IndexcovCohort replaces Index.init + NormalizedDepth for every sample at once
// (indexcov/indexcov.go:83-151): sizes in CSR layout by sample.
func (c *Ctx) IndexcovCohort(sizes []int64, samplePtr []int64) (medians []float64, depths []float32, err error) {
	S := len(samplePtr) - 1
	medians = make([]float64, S)
	depths = make([]float32, len(sizes))
	rc := C.gl_indexcov_cohort(c.h, (*C.int64_t)(unsafe.Pointer(&sizes[0])), (*C.int64_t)(unsafe.Pointer(&samplePtr[0])),
		C.int32_t(S), (*C.double)(unsafe.Pointer(&medians[0])), (*C.float)(unsafe.Pointer(&depths[0])))
	return medians, depths, c.err(rc)
}
