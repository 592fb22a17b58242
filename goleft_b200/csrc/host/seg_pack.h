// seg_pack.h — shared by the int32 packer (seg_pack.cpp) and the BAM feeder (bam_feed.cpp): sorted piece lists ->
// the engine's packed8 words (or globally sorted int32 arrays), assembled in parallel.
#pragma once
#include <stdint.h>
#include <utility>
#include <vector>

namespace glhost {

typedef std::pair<int32_t, int32_t> Piece;          // (start, length >= 1)

struct P8Out { std::vector<int32_t> anchors; std::vector<uint8_t> ds, len; int64_t nb = 0; };

void sort_nearly_sorted(Piece* seg, size_t m);      // bounded insertion sort, std::stable_sort when the disorder is not local
// pieces longer than 255 are cut into 255-base pieces and the list is put back in start order
void split_long_pieces(std::vector<Piece>& list);
// blocks of 64 slots from pieces (length <= 255) in start order
void encode_pieces(const Piece* seg, size_t m, P8Out& o);

// P lists, each sorted by start; the lists are in position order up to a little overlap (list k may hold a few pieces
// that start beyond the first piece of list k+1: second blocks of deletion reads at a unit edge).  With
// x_k = min start over lists >= k, task k takes the pieces of every list j <= k that start in [x_k, x_{k+1}).
// assemble_p8: packed8 blocks with globally sorted anchors; assemble_i32: globally sorted (start, end) arrays.
struct Assembled {
    std::vector<P8Out> parts;                       // per task (assemble_p8)
    std::vector<std::vector<Piece>> merged;         // per task (assemble_i32, or scratch)
    std::vector<int64_t> off;                       // [P+1] block / segment offsets of the tasks
    int64_t total = 0;
};
void assemble_p8(const std::vector<const std::vector<Piece>*>& lists, int threads, Assembled& a);
void assemble_i32(const std::vector<const std::vector<Piece>*>& lists, int threads, Assembled& a);
// copy the assembled parts to contiguous destinations (parallel); dstart/len hold 64 bytes per block
void concat_p8(const Assembled& a, int threads, int32_t* anchors, uint8_t* dstart, uint8_t* len);
void concat_i32(const Assembled& a, int threads, int32_t* start, int32_t* end);

}  // namespace glhost
