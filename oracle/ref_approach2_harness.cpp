// TEST INFRASTRUCTURE — not part of the product path.
//
// The reference's stand-alone CPU Barnes-Hut program (implementation/main_approach_2.cpp: the same PR quadtree,
// ComputeMass and theta-traversal as project.cu's CPU path, WITHOUT the depth cap and the text loader; BASELINE
// config 1 names the main_approach CPU programs), compiled where it lies by textual inclusion
// (-DREF_APPROACH2_SOURCE=...; only `main` is renamed).  N_BODIES = 1000 is hard-coded there
// (main_approach_2.cpp:14).  Drives its own buildTree (:243-259) and computeForces (:261-343) on given bodies and
// dumps the node table and the forces as raw FP64, so that tests can check that oracle/bh_oracle.c (the
// restatement of project.cu) is bit-identical to it whenever the cap is not reached.
// usage: ref_approach2 bodies.bin out.bin     bodies.bin: u64 N | mass[N] | pos[2N] | vel[2N]
//        out.bin: u64 n_nodes | nodes[n_nodes][12] | forces[N][2]
#define main bh_reference_main_unused
#include REF_APPROACH2_SOURCE
#undef main

#include <cstdint>
#include <cstdio>
#include <memory>

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: ref_approach2 bodies.bin out.bin\n"); return 2; }
    auto masses = std::make_unique<Masses>();
    auto positions = std::make_unique<Positions>();
    auto velocities = std::make_unique<Velocities>();
    auto forces = std::make_unique<Forces>();
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) { perror(argv[1]); return 1; }
    uint64_t n = 0;
    if (fread(&n, 8, 1, fi) != 1 || n != (uint64_t)N_BODIES) { fprintf(stderr, "need exactly %d bodies\n", N_BODIES); return 1; }
    if (fread(masses->data(), 8, n, fi) != n || fread(positions->data(), 8, 2 * n, fi) != 2 * n ||
        fread(velocities->data(), 8, 2 * n, fi) != 2 * n) { fprintf(stderr, "short input\n"); return 1; }
    fclose(fi);
    quadtree = buildTree(*positions, *masses);                    // main_approach_2.cpp:243-259
    computeForces(*positions, *masses, *forces);                  // main_approach_2.cpp:261-343
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) { perror(argv[2]); return 1; }
    uint64_t nn = quadtree.size();
    fwrite(&nn, 8, 1, fo);
    fwrite(quadtree[0].data(), 8, nn * 12, fo);
    fwrite((*forces)[0].data(), 8, 2 * n, fo);
    fclose(fo);
    return 0;
}
