// Exercises include/zoe_cuda.hpp (the C++ mirror of zoe's interface) against zoe's doc-test known answers.
// Built by tests/test_cpp_mirror.py with g++; run on the GPU box only.
#include <cstdio>
#include <cstdlib>

#include "zoe_cuda.hpp"

#define CHECK(c)                                                  \
    do {                                                          \
        if (!(c)) {                                               \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                             \
        }                                                         \
    } while (0)

int main() {
    using namespace zoe::cuda;
    // validation happens before any device work (profile.rs:32-44)
    try {
        CudaProfiles::new_with_w256({"ACGT"}, WeightMatrix::new_dna_matrix(2, -5), -1, -10);
        return 1;
    } catch (const ProfileError &e) {
        CHECK(e.kind == ProfileError::BadGapWeights);
    }
    // doc striped.rs:418-441 / sw/mod.rs:164-188: profile = query, SeqSrc::Reference(reference)
    auto w = WeightMatrix::new_dna_matrix(4, -2);
    auto prof = CudaProfiles::new_with_w256({"CGTTCGCCATAAAGGGGG", "CTCAGATTG"}, w, -3, -1, SeqSrc::Reference);
    auto s = prof.sw_score_batch({"ATGCATCGATCGATCGATCGATCGATCGATGC", "GGCCACAGGATTGAG"});
    CHECK(s[0].unwrap() == 26 && s[3].unwrap() == 27);
    auto a = prof.sw_align_batch({"ATGCATCGATCGATCGATCGATCGATCGATGC", "GGCCACAGGATTGAG"});
    CHECK(a[0].unwrap().cigar() == "6M2D9M3S" && a[0].unwrap().score == 26);
    CHECK(a[0].unwrap().query_range.first == 0 && a[0].unwrap().query_range.second == 15);
    CHECK(a[0].unwrap().ref_range.first == 14 && a[0].unwrap().ref_range.second == 31);
    CHECK(a[3].unwrap().cigar() == "5M1D4M" && a[3].unwrap().ref_range.first == 3);
    // profile_set.rs:192-208: sw_align_from_i8_3pass(SeqSrc::Reference(reference)).score == 26
    auto a3 = prof.sw_align_3pass_batch({"ATGCATCGATCGATCGATCGATCGATCGATGC", "GGCCACAGGATTGAG"});
    CHECK(a3[0].unwrap().score == 26 && a3[0].unwrap().ref_range.first == 14 && a3[0].unwrap().ref_range.second == 31);
    CHECK(a3[3].unwrap().score == 27 && a3[3].unwrap().cigar() == "5M1D4M");
    // doc profile_set.rs:293-310: sw_score_ranges_from_i8 -> score 26, query 0..15, ref 14..31
    auto r = prof.sw_score_ranges_batch({"ATGCATCGATCGATCGATCGATCGATCGATGC"});
    CHECK(r[0].unwrap().score == 26 && r[0].unwrap().query_range.second == 15 && r[0].unwrap().ref_range.first == 14);
    prof.set_memory_budget(1 << 20);
    CHECK(prof.sw_align_batch({"ATGCATCGATCGATCGATCGATCGATCGATGC"})[0].unwrap().cigar() == "6M2D9M3S");
    // doc sneaky_snake.rs:54-62
    auto sn = sneaky_snake_batch({"GGTGCAGAGCTC", "AAAA"}, {"GGTGAGAGTTGT", "AAAAAAAA"}, 0.25f);
    CHECK(sn[0].has_value() && *sn[0] == true && !sn[1].has_value());
    // sw/test.rs:81-84 and an Unmapped pair
    auto w25 = WeightMatrix::new_dna_matrix(2, -5);
    auto p2 = CudaProfiles::new_with_w256({std::string(100, 'A')}, w25, -10, -1);
    auto s2 = p2.sw_score_batch({std::string(100, 'A'), "CCCC"});
    CHECK(s2[0].unwrap() == 200 && s2[1].status == Status::Unmapped);
    std::puts("cpp mirror ok");
    return 0;
}
