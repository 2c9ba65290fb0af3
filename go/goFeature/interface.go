package goFeature

// The reference's four interfaces (interface.go:12-143) with the two leaked third-party
// types replaced: Block.Accquire loses its *cublas.Handle and worker arguments' meaning (they
// are accepted and ignored), NewCache takes this package's *Context instead of *cuda.Context.

// Cache : interface of cache, the main object of features (interface.go:12-34)
type Cache interface {
	NewSet(name string, dims int, precision int, batch int) error
	DestroySet(name string) error
	GetSet(name string) (Set, error)
	GetBlockSize() int
	GetEmptyBlock(blocknum int) ([]Block, error)
}

// Set : interface of set (interface.go:37-66)
type Set interface {
	Add(features ...Feature) error
	Delete(ids ...FeatureID) (deleted []FeatureID, err error)
	Update(features ...Feature) (updated []FeatureID, err error)
	Read(ids ...FeatureID) (features []Feature, err error)
	Destroy() error
	Search(threshold FeatureScore, limit int, features ...FeatureValue) (ret [][]FeatureSearchResult, err error)
}

// Block : interface of block, the basic scheduling unit (interface.go:69-117)
type Block interface {
	Capacity() int
	Margin() int
	IsOwned() bool
	Accquire(handle interface{}, owner string, dims int, premision int, batch int, worker interface{}) error
	Release() error
	Insert(features ...Feature) error
	Delete(...FeatureID) (deleted []FeatureID, err error)
	Update(features ...Feature) (updated []FeatureID, err error)
	Read(ids ...FeatureID) (features []Feature, err error)
	Search(inputBuffer, outputBuffer Buffer, batch int, limit int) (ret [][]FeatureSearchResult, err error)
}

// Buffer : device memory buffer (interface.go:120-143)
type Buffer interface {
	GetBuffer() interface{}
	Write(value FeatureValue) error
	Read() (value FeatureValue, err error)
	Copy(src Buffer) error
	Reset() error
	Slice(start, end int) (Buffer, error)
	Size() int
}
