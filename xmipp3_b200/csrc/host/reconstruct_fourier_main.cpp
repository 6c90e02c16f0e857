// xmipp_reconstruct_fourier_b200 — CLI entry, the shape of every Xmipp program main
// (applications/programs/reconstruct_fourier/reconstruct_fourier_main.cpp:27-33):
//   Prog p; p.read(argc, argv); return p.tryRun();
#include <iostream>

#include "prog_rec_fourier.h"

int main(int argc, char** argv) {
    rfhost::ProgRecFourierB200 prog;
    try {
        prog.readParams(argc, argv);
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 2;
    }
    return prog.tryRun();
}
