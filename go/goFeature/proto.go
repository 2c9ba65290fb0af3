package goFeature

// Wire types of the search path.  Same names, same meaning as the reference's proto.go:3-26
// (five declarations; this file carries no build tag, exactly like the reference's).

// FeatureValue : bytes in little endian
type FeatureValue []byte

// FeatureID : string id
type FeatureID string

// FeatureScore : match score, float32
type FeatureScore float32

// Feature : one stored vector (dimension * precision bytes) and its unique id
type Feature struct {
	Value FeatureValue
	ID    FeatureID
}

// FeatureSearchResult : one hit of a search
type FeatureSearchResult struct {
	Score FeatureScore
	ID    FeatureID
}
