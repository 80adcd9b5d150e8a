// TEST INFRASTRUCTURE ONLY.  C entry point around the reference's own StripedSmithWaterman::Aligner
// (helen/modules/src/local_reassembly/ssw_cpp.cpp:320, Align_cpp), called exactly the way
// helen/modules/python/Stitch.py:110-135 calls it (default Filter, maskLen 0).  The reference
// sources are compiled where they lie under /root/reference by build_ref.py; nothing is copied.
#include <cstdint>
#include <cstring>
#include <string>
#include "local_reassembly/ssw_cpp.h"

extern "C" {

struct ref_ssw_result {
    int32_t score, score2, ref_begin, ref_end, query_begin, query_end, ref_end2, mismatches, cigar_len;
};

// returns 0 on success, 1 if Align_cpp refused the input, 2 if the cigar buffer is too small
int ref_ssw_align(const char* ref, int32_t ref_len, const char* query,
                  int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                  ref_ssw_result* out, char* cigar, int32_t cigar_cap) {
    StripedSmithWaterman::Aligner aligner((uint8_t)match, (uint8_t)mismatch, (uint8_t)gap_open, (uint8_t)gap_extend);
    StripedSmithWaterman::Filter filter;
    StripedSmithWaterman::Alignment al;
    aligner.SetReferenceSequence(ref, ref_len);
    if (!aligner.Align_cpp(query, filter, &al, 0)) return 1;
    out->score = al.sw_score;
    out->score2 = al.sw_score_next_best;
    out->ref_begin = al.ref_begin;
    out->ref_end = al.ref_end;
    out->query_begin = al.query_begin;
    out->query_end = al.query_end;
    out->ref_end2 = al.ref_end_next_best;
    out->mismatches = al.mismatches;
    out->cigar_len = (int32_t)al.cigar_string.size();
    if ((int32_t)al.cigar_string.size() + 1 > cigar_cap) return 2;
    std::memcpy(cigar, al.cigar_string.c_str(), al.cigar_string.size() + 1);
    return 0;
}

}
