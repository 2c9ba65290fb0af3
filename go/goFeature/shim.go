package goFeature

// cgo binding of libgofeature_b200.so (include/gofeature_b200.h).  It stands where the
// reference imports github.com/unixpickle/cuda and .../cuda/cublas (block.go:10-11,
// buffer.go:5, cache.go:6-7, set.go:9-10).  No logic lives here: every method packs its
// arguments into flat C buffers (cgo may not pass Go pointers to Go pointers), makes ONE call,
// and maps the status code onto the error.go sentinels.
//
// NOT COMPILED in the build container (no Go toolchain there or on the GPU box); the behaviour
// behind each call is tested through the same C ABI from tests/ (ctypes).

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gofeature_b200 -lgofeature_b200 -Wl,-rpath,${SRCDIR}/../../gofeature_b200
#include <stdlib.h>
#include "gofeature_b200.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"
)

// errFromStatus maps a status code onto the error.go sentinels.  Codes >= 100 carry a detail string that the
// library keeps PER OS THREAD (gf_last_error_detail): a goroutine may migrate between the failing call and the
// fetch, so callers that can fail with such a code wrap both in withThread (runtime.LockOSThread).
func errFromStatus(st C.int) error {
	if st == 0 {
		return nil
	}
	if int(st) < len(statusTable) {
		return statusTable[st]
	}
	return errors.New(C.GoString(C.gf_strerror(st)) + ": " + C.GoString(C.gf_last_error_detail()))
}

// withThread runs one C call and the translation of its status on the same OS thread.
func withThread(call func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	return errFromStatus(call())
}

// Context replaces *cuda.Context (search_test.go:40: cuda.NewContext(devices[0], -1)).
type Context struct{ h *C.gf_context }

// AllDevices replaces cuda.AllDevices (search_test.go:33): device ordinals.
func AllDevices() ([]int, error) {
	var n C.int
	if err := errFromStatus(C.gf_device_count(&n)); err != nil {
		return nil, err
	}
	out := make([]int, int(n))
	for i := range out {
		out[i] = i
	}
	return out, nil
}

// NewContext replaces cuda.NewContext.
func NewContext(device int, _ int) (*Context, error) {
	c := &Context{}
	if err := withThread(func() C.int { return C.gf_context_create(C.int(device), &c.h) }); err != nil {
		return nil, err
	}
	return c, nil
}

// NewContextMulti : one context over several devices of this process (gf_context_create_multi).  A cache created
// on it stripes its blocks over the devices (block i on devices[i % n], the order GetEmptyBlock hands them out,
// cache.go:104-115), Set.Search scans all of them and returns one merged, id-resolved result; everything else of
// the Cache / Set API is unchanged.  The reference has no counterpart (cache.go:21-43 builds on one context).
func NewContextMulti(devices []int) (*Context, error) {
	if len(devices) == 0 {
		return nil, ErrInvalidDeviceID
	}
	ds := make([]C.int, len(devices))
	for i, d := range devices {
		ds[i] = C.int(d)
	}
	c := &Context{}
	if err := withThread(func() C.int { return C.gf_context_create_multi(&ds[0], C.int(len(ds)), &c.h) }); err != nil {
		return nil, err
	}
	return c, nil
}

// TotalMem replaces Device.TotalMem (demo/main.go:61).
func TotalMem(device int) (uint64, error) {
	var b C.uint64_t
	err := errFromStatus(C.gf_device_total_mem(C.int(device), &b))
	return uint64(b), err
}

func packIDs(ids []FeatureID) (unsafe.Pointer, *C.uint64_t, func()) {
	total := 0
	for _, id := range ids {
		total += len(id)
	}
	buf := C.malloc(C.size_t(total + 1))
	off := (*[1 << 28]C.uint64_t)(C.malloc(C.size_t(8 * (len(ids) + 1))))
	pos := 0
	for i, id := range ids {
		off[i] = C.uint64_t(pos)
		copy((*[1 << 30]byte)(buf)[pos:pos+len(id)], id)
		pos += len(id)
	}
	off[len(ids)] = C.uint64_t(pos)
	return buf, &off[0], func() { C.free(buf); C.free(unsafe.Pointer(off)) }
}

type cache struct {
	h   *C.gf_cache
	ctx *Context
}

// NewCache : cache.go:21-43
func NewCache(ctx *Context, blockNum, blockSize int) (Cache, error) {
	c := &cache{ctx: ctx}
	if err := errFromStatus(C.gf_cache_create(ctx.h, C.int(blockNum), C.int64_t(blockSize), &c.h)); err != nil {
		return nil, err
	}
	return c, nil
}

// NewCacheNoShadow : NewCache without the shadow slab (by default int8 rows + one float scale per row: blockSize/4 +
// blockSize/128 more bytes per block; the tensor path then streams the float32 rows).  Results are identical either way.
func NewCacheNoShadow(ctx *Context, blockNum, blockSize int) (Cache, error) {
	c := &cache{ctx: ctx}
	if err := errFromStatus(C.gf_cache_create_ex(ctx.h, C.int(blockNum), C.int64_t(blockSize),
		C.uint(C.GF_CACHE_NO_SHADOW), &c.h)); err != nil {
		return nil, err
	}
	return c, nil
}

func (c *cache) NewSet(name string, dims, precision, batch int) error {
	n := C.CString(name)
	defer C.free(unsafe.Pointer(n))
	return errFromStatus(C.gf_cache_new_set(c.h, n, C.int(dims), C.int(precision), C.int(batch)))
}

func (c *cache) DestroySet(name string) error {
	n := C.CString(name)
	defer C.free(unsafe.Pointer(n))
	return errFromStatus(C.gf_cache_destroy_set(c.h, n))
}

func (c *cache) GetSet(name string) (Set, error) {
	n := C.CString(name)
	defer C.free(unsafe.Pointer(n))
	s := &set{}
	if err := errFromStatus(C.gf_cache_get_set(c.h, n, &s.h)); err != nil {
		return nil, err
	}
	return s, nil
}

func (c *cache) GetBlockSize() int { return int(C.gf_cache_get_block_size(c.h)) }

func (c *cache) GetEmptyBlock(blocknum int) ([]Block, error) {
	hs := make([]*C.gf_block, blocknum+1)
	if err := errFromStatus(C.gf_cache_get_empty_block(c.h, C.int(blocknum), &hs[0])); err != nil {
		return nil, err
	}
	out := make([]Block, blocknum)
	for i := range out {
		out[i] = &block{h: hs[i]}
	}
	return out, nil
}

type set struct{ h *C.gf_set }

func (s *set) Add(features ...Feature) error {
	if len(features) == 0 {
		return nil
	}
	value := FeatureToFeatureValue1D(features...) // util.go:14-19
	ids := make([]FeatureID, len(features))
	for i, f := range features {
		ids[i] = f.ID
	}
	buf, off, free := packIDs(ids)
	defer free()
	return errFromStatus(C.gf_set_add(s.h, C.int64_t(len(features)), unsafe.Pointer(&value[0]),
		C.uint64_t(len(value)), (*C.char)(buf), off))
}

func (s *set) Delete(ids ...FeatureID) (deleted []FeatureID, err error) {
	if len(ids) == 0 {
		return nil, nil
	}
	buf, off, free := packIDs(ids)
	defer free()
	flags := make([]C.uint8_t, len(ids))
	var n C.int64_t
	if err = errFromStatus(C.gf_set_delete(s.h, C.int64_t(len(ids)), (*C.char)(buf), off, &flags[0], &n)); err != nil {
		return nil, err
	}
	for i, f := range flags {
		if f != 0 {
			deleted = append(deleted, ids[i])
		}
	}
	return
}

func (s *set) Update(features ...Feature) (updated []FeatureID, err error) {
	if len(features) == 0 {
		return nil, nil
	}
	value := FeatureToFeatureValue1D(features...)
	ids := make([]FeatureID, len(features))
	for i, f := range features {
		ids[i] = f.ID
	}
	buf, off, free := packIDs(ids)
	defer free()
	flags := make([]C.uint8_t, len(ids))
	var n C.int64_t
	if err = errFromStatus(C.gf_set_update(s.h, C.int64_t(len(ids)), unsafe.Pointer(&value[0]),
		C.uint64_t(len(value)), (*C.char)(buf), off, &flags[0], &n)); err != nil {
		return nil, err
	}
	for i, f := range flags {
		if f != 0 {
			updated = append(updated, ids[i])
		}
	}
	return
}

func (s *set) Read(ids ...FeatureID) (features []Feature, err error) {
	if len(ids) == 0 {
		return nil, nil
	}
	rowb := int(C.gf_set_dims(s.h)) * int(C.gf_set_precision(s.h))
	buf, off, free := packIDs(ids)
	defer free()
	out := make([]byte, len(ids)*rowb)
	flags := make([]C.uint8_t, len(ids))
	var n C.int64_t
	if err = errFromStatus(C.gf_set_read(s.h, C.int64_t(len(ids)), (*C.char)(buf), off, unsafe.Pointer(&out[0]),
		C.uint64_t(len(out)), &flags[0], &n)); err != nil {
		return nil, err
	}
	for i, f := range flags {
		if f != 0 {
			features = append(features, Feature{Value: out[i*rowb : (i+1)*rowb], ID: ids[i]})
		}
	}
	return
}

func (s *set) Destroy() error { return errFromStatus(C.gf_set_destroy(s.h)) }

// Search : set.go:118-159.  One cgo crossing for the search, one for the export.
func (s *set) Search(threshold FeatureScore, limit int, features ...FeatureValue) (ret [][]FeatureSearchResult, err error) {
	var flat FeatureValue
	for _, f := range features {
		flat = append(flat, f...)
	}
	var p unsafe.Pointer
	if len(flat) > 0 {
		p = unsafe.Pointer(&flat[0])
	}
	var r *C.gf_result
	if err = errFromStatus(C.gf_set_search(s.h, C.float(threshold), C.int(limit), C.int(len(features)), p,
		C.uint64_t(len(flat)), &r)); err != nil {
		return nil, err
	}
	defer C.gf_result_free(r)
	return exportResult(r), nil
}

func exportResult(r *C.gf_result) [][]FeatureSearchResult {
	q := int(C.gf_result_batch(r))
	total := int(C.gf_result_total(r))
	counts := make([]C.int32_t, q+1)
	scores := make([]C.float, total+1)
	ids := make([]byte, int(C.gf_result_id_bytes(r))+1)
	offs := make([]C.uint64_t, total+1)
	C.gf_result_export(r, &counts[0], &scores[0], nil, (*C.char)(unsafe.Pointer(&ids[0])), &offs[0])
	ret := make([][]FeatureSearchResult, q)
	k := 0
	for i := 0; i < q; i++ {
		for j := 0; j < int(counts[i]); j++ {
			ret[i] = append(ret[i], FeatureSearchResult{Score: FeatureScore(scores[k]),
				ID: FeatureID(ids[offs[k]:offs[k+1]])})
			k++
		}
	}
	return ret
}

type block struct{ h *C.gf_block }

func (b *block) Capacity() int { return int(C.gf_block_capacity(b.h)) }
func (b *block) Margin() int   { return int(C.gf_block_margin(b.h)) }
func (b *block) IsOwned() bool { return C.gf_block_is_owned(b.h) != 0 }
func (b *block) Accquire(_ interface{}, owner string, dims, precision, batch int, _ interface{}) error {
	o := C.CString(owner)
	defer C.free(unsafe.Pointer(o))
	return errFromStatus(C.gf_block_acquire(b.h, o, C.int(dims), C.int(precision), C.int(batch)))
}
func (b *block) Release() error { return errFromStatus(C.gf_block_release(b.h)) }
func (b *block) Insert(features ...Feature) error {
	if len(features) == 0 {
		return nil
	}
	value := FeatureToFeatureValue1D(features...)
	ids := make([]FeatureID, len(features))
	for i, f := range features {
		ids[i] = f.ID
	}
	buf, off, free := packIDs(ids)
	defer free()
	return errFromStatus(C.gf_block_insert(b.h, C.int64_t(len(features)), unsafe.Pointer(&value[0]),
		C.uint64_t(len(value)), (*C.char)(buf), off))
}
func (b *block) Delete(ids ...FeatureID) (deleted []FeatureID, err error) {
	if len(ids) == 0 {
		return nil, nil
	}
	buf, off, free := packIDs(ids)
	defer free()
	flags := make([]C.uint8_t, len(ids))
	var n C.int64_t
	if err = errFromStatus(C.gf_block_delete(b.h, C.int64_t(len(ids)), (*C.char)(buf), off, &flags[0], &n)); err != nil {
		return nil, err
	}
	for i, f := range flags {
		if f != 0 {
			deleted = append(deleted, ids[i])
		}
	}
	return
}
func (b *block) Update(features ...Feature) ([]FeatureID, error) { return nil, nil } // block.go:169-172
func (b *block) Read(ids ...FeatureID) ([]Feature, error)        { return nil, nil } // block.go:176-179
func (b *block) Search(inputBuffer, outputBuffer Buffer, batch, limit int) ([][]FeatureSearchResult, error) {
	var r *C.gf_result
	in := inputBuffer.(*gpuBuffer)
	var out *C.gf_buffer
	if ob, ok := outputBuffer.(*gpuBuffer); ok {
		out = ob.h
	}
	if err := errFromStatus(C.gf_block_search(b.h, in.h, out, C.int(batch), C.int(limit), &r)); err != nil {
		return nil, err
	}
	defer C.gf_result_free(r)
	ret := exportResult(r)
	if len(ret) == 0 {
		return nil, nil // block.go:186-189
	}
	return ret, nil
}

type gpuBuffer struct{ h *C.gf_buffer }

// NewGPUBuffer : buffer.go:15-32 (the allocator argument has no counterpart)
func NewGPUBuffer(ctx *Context, _ interface{}, size int) (Buffer, error) {
	b := &gpuBuffer{}
	if err := errFromStatus(C.gf_buffer_create(ctx.h, C.uint64_t(size), &b.h)); err != nil {
		return nil, err
	}
	return b, nil
}
func (b *gpuBuffer) GetBuffer() interface{} { return unsafe.Pointer(C.gf_buffer_device_ptr(b.h)) }
func (b *gpuBuffer) Write(value FeatureValue) error {
	if len(value) == 0 {
		return nil
	}
	return errFromStatus(C.gf_buffer_write(b.h, unsafe.Pointer(&value[0]), C.uint64_t(len(value))))
}
func (b *gpuBuffer) Read() (FeatureValue, error) {
	v := make(FeatureValue, b.Size()+1)
	err := errFromStatus(C.gf_buffer_read(b.h, unsafe.Pointer(&v[0]), C.uint64_t(b.Size())))
	return v[:b.Size()], err
}
func (b *gpuBuffer) Copy(src Buffer) error { return errFromStatus(C.gf_buffer_copy(b.h, src.(*gpuBuffer).h)) }
func (b *gpuBuffer) Reset() error          { return errFromStatus(C.gf_buffer_reset(b.h)) }
func (b *gpuBuffer) Slice(start, end int) (Buffer, error) {
	s := &gpuBuffer{}
	if err := errFromStatus(C.gf_buffer_slice(b.h, C.int64_t(start), C.int64_t(end), &s.h)); err != nil {
		return nil, err
	}
	return s, nil
}
func (b *gpuBuffer) Size() int { return int(C.gf_buffer_size(b.h)) }

// FeatureToFeatureValue1D : util.go:14-19
func FeatureToFeatureValue1D(features ...Feature) (ret FeatureValue) {
	for _, feature := range features {
		ret = append(ret, feature.Value...)
	}
	return
}

// NoramlizeFloat32 : util.go:125-139, served by the library so Go and the tests share one definition
func NoramlizeFloat32(feature []float32) []float32 {
	out := make([]float32, len(feature))
	if len(feature) > 0 {
		C.gf_util_normalize_float32((*C.float)(unsafe.Pointer(&feature[0])), (*C.float)(unsafe.Pointer(&out[0])),
			C.int64_t(len(feature)))
	}
	return out
}
