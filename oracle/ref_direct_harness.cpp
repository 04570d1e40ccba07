// TEST INFRASTRUCTURE — not part of the product path.
//
// The reference's direct all-pairs force (implementation/main_approach_1.cpp:53-75), compiled where it lies
// (textual inclusion, -DREF_DIRECT_SOURCE=...; only `main` is renamed).  The reference hard-codes n = 2 bodies
// (main_approach_1.cpp:12), so this pins the pair expression of oracle/bh_oracle.c:bho_direct_forces — order of
// operations of distance_squared, sqrt, factor — against the reference's own function, pair by pair.
// usage: ref_direct x0 y0 m0 x1 y1 m1   -> prints the four force components as C99 hex floats
#define main bh_reference_main_unused
#include REF_DIRECT_SOURCE
#undef main

#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
    if (argc != 7) { fprintf(stderr, "usage: ref_direct x0 y0 m0 x1 y1 m1\n"); return 2; }
    Positions positions;
    Masses masses;
    Forces forces;
    positions[0] = {strtod(argv[1], nullptr), strtod(argv[2], nullptr)};
    masses[0] = strtod(argv[3], nullptr);
    positions[1] = {strtod(argv[4], nullptr), strtod(argv[5], nullptr)};
    masses[1] = strtod(argv[6], nullptr);
    computeForces(positions, masses, forces);                     // main_approach_1.cpp:53-75
    printf("%a %a %a %a\n", forces[0][0], forces[0][1], forces[1][0], forces[1][1]);
    return 0;
}
