// TEST INFRASTRUCTURE — not part of the product path.
//
// Harness around the UNMODIFIED reference translation unit.  The reference
// (`implementation/project.cu`) is compiled *where it lies* under
// /root/reference by textual inclusion (path given with -DREF_SOURCE=...);
// nothing of it is copied into this repository.  Only its `main` is renamed
// so that this file can drive the reference's own CPU functions
//
//     buildTree            project.cu:575-591
//     TraverseTreeToFile   project.cu:504-534
//     computeForces        project.cu:593-675
//     updateAccelerations / updateVelocities / updatePositions  project.cu:795-817
//
// in exactly the order runSimulationCpu (project.cu:883-910) calls them, and
// dump every intermediate as raw FP64 so that (a) the C restatement in
// oracle/bh_oracle.c can be pinned bit-for-bit and (b) golden vectors can be
// generated.  Host code only: no CUDA call is reached, so it runs without a GPU.
//
// N_BODIES is a compile-time macro in the reference (std::array sizes), hence
// one binary per N:  oracle/build_ref.sh <N>  ->  oracle/_ref/ref_harness_N<N>
//
// usage: ref_harness_N<N> --in bodies.bin --steps K [--out dump.bin]
//                         [--dump tree,forces,state] [--dump-steps all|first|last]
//                         [--quadtree-txt prefix] [--reset-each-step] [--positions-txt file]
//   --load-text DIR: instead of --in, read the bodies with the reference's own loadSimulationDataFromText
//                    (project.cu:103-161) from DIR/masses_init.txt, positions_init.txt, velocities_init.txt
//   --positions-txt: the reference's trajectory file (savePositions at t = 0 and after every step,
//                    project.cu:855-863, :879, :907, :912), what plot_2d.py reads
//   bodies.bin : u64 N | mass[N] | pos[2N] | vel[2N]      (all FP64, host endian)
//   dump.bin   : records { char name[16]; u64 step; u64 ndoubles; double data[] }
// stdout: one JSON line per step with wall-clock microseconds per phase.
#define main bh_reference_main_unused
#include REF_SOURCE
#undef main

#include <cstdio>
#include <cstring>
#include <cstdint>
#include <memory>
#include <new>
#include <vector>

static void put(FILE* f, const char* name, uint64_t step, const double* d, uint64_t n) {
    if (!f) return;
    char tag[16];
    memset(tag, 0, sizeof tag);
    strncpy(tag, name, 15);
    fwrite(tag, 1, 16, f);
    fwrite(&step, 8, 1, f);
    fwrite(&n, 8, 1, f);
    fwrite(d, 8, n, f);
}

int main(int argc, char** argv) {
    const char* in = nullptr; const char* out = nullptr; const char* qtxt = nullptr; const char* ptxt = nullptr; const char* ldir = nullptr;
    const char* what = "tree,forces,state"; const char* which = "all";
    int steps = 1; bool reset = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--in") && i + 1 < argc) in = argv[++i];
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--dump") && i + 1 < argc) what = argv[++i];
        else if (!strcmp(argv[i], "--dump-steps") && i + 1 < argc) which = argv[++i];
        else if (!strcmp(argv[i], "--quadtree-txt") && i + 1 < argc) qtxt = argv[++i];
        else if (!strcmp(argv[i], "--reset-each-step")) reset = true;
        else if (!strcmp(argv[i], "--positions-txt") && i + 1 < argc) ptxt = argv[++i];
        else if (!strcmp(argv[i], "--load-text") && i + 1 < argc) ldir = argv[++i];
        else { fprintf(stderr, "bad arg %s\n", argv[i]); return 2; }
    }
    if (!in && !ldir) { fprintf(stderr, "--in or --load-text required\n"); return 2; }
    const size_t N = N_BODIES;
    // heap, not stack: the reference keeps these in main's frame (project.cu:1055-1057)
    auto masses = std::make_unique<Masses>();
    // TraverseTreeToFile indexes positions[] with the NEGATIVE encoded occupant of cap-level leaves
    // (project.cu:514-518), i.e. it reads up to (N + 1) entries BELOW the array: keep the array
    // in the upper half of a zero-filled arena so that those reads stay inside mapped memory.
    std::vector<char> arena(2 * sizeof(Positions) + 64, 0);
    static_assert(alignof(Positions) <= 16 && (sizeof(Positions) + 64) % 16 == 0, "arena offset keeps the alignment");
    Positions* positions = new (arena.data() + sizeof(Positions) + 64) Positions();
    auto velocities = std::make_unique<Velocities>();
    auto accelerations = std::make_unique<Accelerations>();
    auto forces = std::make_unique<Forces>();
    auto pos0 = std::make_unique<Positions>();
    auto vel0 = std::make_unique<Velocities>();
    if (ldir) {
        const std::string d(ldir);
        try {                                                                         // project.cu:1065-1066
            loadSimulationDataFromText(d + "/masses_init.txt", d + "/positions_init.txt", d + "/velocities_init.txt",
                                       N_BODIES, *masses, *positions, *velocities);
        } catch (const std::exception& e) { fprintf(stderr, "reference loader: %s\n", e.what()); return 5; }
    } else {
    FILE* fi = fopen(in, "rb");
    if (!fi) { perror(in); return 1; }
    uint64_t n_in = 0;
    if (fread(&n_in, 8, 1, fi) != 1 || n_in != N) {
        fprintf(stderr, "input holds %llu bodies, binary built for %zu\n", (unsigned long long)n_in, N);
        return 1;
    }
    if (fread(masses->data(), 8, N, fi) != N || fread(positions->data(), 8, 2 * N, fi) != 2 * N ||
        fread(velocities->data(), 8, 2 * N, fi) != 2 * N) { fprintf(stderr, "short input\n"); return 1; }
    fclose(fi);
    }
    *pos0 = *positions; *vel0 = *velocities;
    FILE* fo = out ? fopen(out, "wb") : nullptr;
    const bool d_tree = strstr(what, "tree"), d_forces = strstr(what, "forces"), d_state = strstr(what, "state");

    if (fo && ldir) {   // what the reference's loader produced
        put(fo, "loaded_mass", 0, masses->data(), N);
        put(fo, "loaded_pos", 0, (*positions)[0].data(), 2 * N);
        put(fo, "loaded_vel", 0, (*velocities)[0].data(), 2 * N);
    }
    std::string output_str;
    double absolute_t = 0.0;
    if (ptxt) savePositions(output_str, *positions, absolute_t);                      // project.cu:879
    for (int step = 0; step < steps; ++step) {
        absolute_t += DELTA_T;                                                       // project.cu:884
        if (reset) { *positions = *pos0; *velocities = *vel0; }
        const bool dump = !strcmp(which, "all") || (!strcmp(which, "first") && step == 0) ||
                          (!strcmp(which, "last") && step == steps - 1);
        auto t0 = std::chrono::high_resolution_clock::now();
        quadtree = buildTree(*positions, *masses);                                   // project.cu:887
        auto t1 = std::chrono::high_resolution_clock::now();
        if (qtxt && (step == 0 || step == steps - 1)) {                              // project.cu:890-893
            std::string name = std::string(qtxt) + (step == 0 ? "_init.txt" : "_final.txt");
            if (step == 0 || steps > 1) {
                std::ofstream tf(name);
                TraverseTreeToFile(0, tf, *positions);
            }
        }
        if (dump && d_tree) put(fo, "tree", step, quadtree[0].data(), quadtree.size() * QUADRANT_SIZE);
        auto t2 = std::chrono::high_resolution_clock::now();
        computeForces(*positions, *masses, *forces);                                 // project.cu:897
        auto t3 = std::chrono::high_resolution_clock::now();
        updateAccelerations(*forces, *masses, *accelerations);                       // project.cu:899
        updateVelocities(*velocities, *accelerations, DELTA_T);                      // project.cu:901
        updatePositions(*positions, *velocities, DELTA_T);                           // project.cu:903
        auto t4 = std::chrono::high_resolution_clock::now();
        if (dump && d_forces) put(fo, "forces", step, (*forces)[0].data(), 2 * N);
        if (dump && d_state) {
            put(fo, "acc", step, (*accelerations)[0].data(), 2 * N);
            put(fo, "vel", step, (*velocities)[0].data(), 2 * N);
            put(fo, "pos", step, (*positions)[0].data(), 2 * N);
        }
        if (ptxt) savePositions(output_str, *positions, absolute_t);                  // project.cu:907
        auto us = [](auto a, auto b) { return (long long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
        printf("{\"step\": %d, \"n_bodies\": %zu, \"nodes\": %zu, \"build_us\": %lld, \"force_us\": %lld, \"update_us\": %lld}\n",
               step, N, quadtree.size(), us(t0, t1), us(t2, t3), us(t3, t4));
        fflush(stdout);
    }
    if (fo) fclose(fo);
    if (ptxt) { std::ofstream pf(ptxt); pf << output_str; }                           // project.cu:912
    return 0;
}
