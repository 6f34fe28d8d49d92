// fast_deflate.hh -- the sink's compressor: one raw DEFLATE (RFC 1951) stream per BGZF block.
//
// The reference writes through htslib's bgzf_write (htsiter.cc:67-86), i.e. zlib or libdeflate at the default level.
// Recalibrated FASTQ / BAM is mostly near-random bases and low-entropy qualities: nearly all of what a general
// compressor gains on it comes from entropy coding, little from string matching (the read names are the exception), and
// zlib's short matches even cost bits on the quality lines.  This encoder therefore does one pass of greedy matching
// that keeps only LONG matches (names, adapters, duplicated reads), one dynamic Huffman code per block, and emits the
// block; it never reads a byte twice outside the L1.  Any inflater reads the result.
#pragma once
#include <cstddef>
#include <cstdint>

namespace kbbq {
namespace fdeflate {

// most input bytes one call takes (a stored block's limit; BGZF blocks carry 0xff00)
static const size_t kMaxInput = 65535;
// bytes `compress` may need for n input bytes in the worst case (it falls back to a stored block)
inline size_t bound(size_t n) { return n + 16; }
// Compresses in[0, n), n <= kMaxInput, as ONE final DEFLATE block into out[0, cap); returns the stream's size in bytes,
// 0 when cap < bound(n) or n > kMaxInput.  min_match: shortest match kept (>= 8; > 258 = no matching, Huffman only).
size_t compress(const uint8_t* in, size_t n, uint8_t* out, size_t cap, int min_match);

}  // namespace fdeflate
}  // namespace kbbq
