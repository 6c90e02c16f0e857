// integration/reconstruct_fourier_b200_adapter.cpp — the -DWITH_XMIPPCORE seam.
//
// Compiled ONLY inside an Xmipp build (xmippCore on the include path, -DWITH_XMIPPCORE): a program class that derives
// from the reference's ProgReconsBase (libraries/reconstruction/recons.h:36-44: setIO + run), keeps xmippCore's
// MetaDataVec / Image<T> / SymList / CTFDescription for everything on the host side and hands the particles to the
// B200 library through the C ABI of include/recfourier_b200.h.  It replaces the body of ProgRecFourierGPU
// (libraries/reconstruction_adapt_cuda/reconstruct_fourier_gpu.cpp: processImages :832-872, finishComputations :879-932).
// Without WITH_XMIPPCORE this file is empty: xmippCore is not part of the reference tree, so the
// self-contained host program of xmipp3_b200/csrc/host/ is what this repository builds and tests.
#ifdef WITH_XMIPPCORE

#include <core/metadata_vec.h>
#include <core/symmetries.h>
#include <core/xmipp_image.h>
#include <core/xmipp_program.h>
#include <data/ctf.h>
#include <reconstruction/recons.h>

#include <vector>

#include "recfourier_b200.h"

class ProgRecFourierB200 : public ProgReconsBase {
public:
    FileName fn_in, fn_out, fn_sym;
    double padding_factor_proj = 2, padding_factor_vol = 2, maxResolution = 0.5, minCTF = 0.01, Ts = 1;
    double blob_radius = 1.9, blob_alpha = 15;
    int blob_order = 0, NiterWeight = 1, device = 0, bufferSize = 1024;
    bool do_weights = false, useCTF = false, phaseFlipped = false, fast = false;

    void defineParams() override {      // the flag set of reconstruct_fourier.cpp:42-58 + reconstruct_fourier_gpu.cpp:48-72
        addParamsLine("   -i <md_file>                : Metadata file with input projections");
        addParamsLine("  [-o <volume_file=\"rec_fourier.vol\">]  : Filename for output volume");
        addParamsLine("  [--iter <iterations=1>]      : Number of iterations for weight correction");
        addParamsLine("  [--sym <symfile=c1>]         : Enforce symmetry in projections");
        addParamsLine("  [--padding <proj=2.0> <vol=2.0>] : Padding used for projections and volume");
        addParamsLine("  [--max_resolution <p=0.5>]   : Max resolution (Nyquist=0.5)");
        addParamsLine("  [--weight]                   : Use weights stored in the image metadata");
        addParamsLine("  [--blob <radius=1.9> <order=0> <alpha=15>] : Blob parameters");
        addParamsLine("  [--useCTF]                   : Use CTF information if present");
        addParamsLine("  [--sampling <Ts=1>]          : sampling rate of the input images in Angstroms/pixel");
        addParamsLine("  [--phaseFlipped]             : Give this flag if images have been already phase flipped");
        addParamsLine("  [--minCTF <ctf=0.01>]        : Minimum value of the CTF that will be inverted");
        addParamsLine("  [--device <dev=0>]           : GPU device to use");
        addParamsLine("  [--bufferSize <size=1024>]   : Number of projections handed to the GPU per call");
        addParamsLine("  [--fast]                     : Do the blobing at the end of the computation");
    }
    void readParams() override {
        fn_in = getParam("-i");  fn_out = getParam("-o");  fn_sym = getParam("--sym");
        NiterWeight = getIntParam("--iter");
        padding_factor_proj = getDoubleParam("--padding", 0);  padding_factor_vol = getDoubleParam("--padding", 1);
        maxResolution = getDoubleParam("--max_resolution");
        do_weights = checkParam("--weight");
        blob_radius = getDoubleParam("--blob", 0);  blob_order = getIntParam("--blob", 1);  blob_alpha = getDoubleParam("--blob", 2);
        useCTF = checkParam("--useCTF");  phaseFlipped = checkParam("--phaseFlipped");
        minCTF = getDoubleParam("--minCTF");  if (useCTF) Ts = getDoubleParam("--sampling");
        device = getIntParam("--device");  bufferSize = getIntParam("--bufferSize");  fast = checkParam("--fast");
    }
    void setIO(const FileName& in, const FileName& out) override { fn_in = in;  fn_out = out; }

    void run() override {
        MetaDataVec SF;
        SF.read(fn_in);
        SF.removeDisabled();
        size_t Xdim, Ydim, Zdim, Ndim;
        getImageSize(SF, Xdim, Ydim, Zdim, Ndim);
        if (Xdim != Ydim) REPORT_ERROR(ERR_MULTIDIM_SIZE, "This algorithm only works for squared images");
        // R_repository without the identity (reconstruct_fourier.cpp:272-286), row-major 3x3
        std::vector<double> sym;
        SymList SL;
        if (fn_sym != "") SL.readSymmetryFile(fn_sym);
        for (int isym = 0; isym < SL.symsNo(); isym++) {
            Matrix2D<double> L(4, 4), R(4, 4);
            SL.getMatrices(isym, L, R);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) sym.push_back(MAT_ELEM(R, i, j));
        }
        rfb200_config cfg = {};
        cfg.abi_version = RFB200_ABI_VERSION;
        cfg.img_size = (int)Xdim;
        cfg.pad_proj = padding_factor_proj;  cfg.pad_vol = padding_factor_vol;
        cfg.max_resolution = maxResolution;
        cfg.blob_radius = blob_radius;  cfg.blob_order = blob_order;  cfg.blob_alpha = blob_alpha;
        cfg.n_sym = (int)(sym.size() / 9);
        cfg.sym_matrices = sym.empty() ? nullptr : sym.data();
        cfg.use_ctf = useCTF && (SF.containsLabel(MDL_CTF_MODEL) || SF.containsLabel(MDL_CTF_DEFOCUSU));
        cfg.phase_flipped = phaseFlipped;  cfg.sampling = Ts;  cfg.min_ctf = minCTF;
        cfg.use_weights = do_weights;  cfg.n_iter_weight = NiterWeight;  cfg.fast = fast;
        cfg.device = device;  cfg.max_batch = bufferSize;
        rfb200_handle h;
        if (rfb200_create(&cfg, &h) != RFB200_OK) REPORT_ERROR(ERR_GPU_MEMORY, rfb200_last_error(nullptr));

        float* images = nullptr;
        rfb200_host_alloc((void**)&images, sizeof(float) * (size_t)bufferSize * Xdim * Xdim);
        std::vector<rfb200_particle> rows(bufferSize);
        size_t n = 0;
        for (size_t objId : SF.ids()) {
            Image<float> img;
            FileName fn;
            SF.getValue(MDL_IMAGE, fn, objId);
            img.read(fn);                                       // the shifts are applied on the GPU (readApplyGeo's job)
            memcpy(images + n * Xdim * Xdim, MULTIDIM_ARRAY(img()), sizeof(float) * Xdim * Xdim);
            rfb200_particle& p = rows[n];
            p = {};
            SF.getValue(MDL_ANGLE_ROT, p.rot, objId);   SF.getValue(MDL_ANGLE_TILT, p.tilt, objId);
            SF.getValue(MDL_ANGLE_PSI, p.psi, objId);   SF.getValue(MDL_SHIFT_X, p.shift_x, objId);
            SF.getValue(MDL_SHIFT_Y, p.shift_y, objId); p.weight = 1;  SF.getValue(MDL_WEIGHT, p.weight, objId);
            p.kV = 100;  p.K = 1;
            if (cfg.use_ctf) {
                CTFDescription ctf;
                ctf.enable_CTF = true;
                ctf.readFromMetadataRow(SF, objId);
                p.kV = ctf.kV;  p.defocusU = ctf.DeltafU;  p.defocusV = ctf.DeltafV;  p.defocus_angle = ctf.azimuthal_angle;
                p.Cs = ctf.Cs;  p.Ca = ctf.Ca;  p.espr = ctf.espr;  p.ispr = ctf.ispr;  p.alpha = ctf.alpha;
                p.DeltaF = ctf.DeltaF;  p.DeltaR = ctf.DeltaR;  p.Q0 = ctf.Q0;  p.K = ctf.K;
                p.envR0 = ctf.envR0;  p.envR1 = ctf.envR1;  p.envR2 = ctf.envR2;
                p.phase_shift = ctf.phase_shift;  p.vpp_radius = ctf.VPP_radius;
            }
            if (++n == (size_t)bufferSize) {
                if (rfb200_insert_batch(h, images, rows.data(), (int)n) != RFB200_OK) REPORT_ERROR(ERR_UNCLASSIFIED, rfb200_last_error(h));
                n = 0;
            }
        }
        if (n && rfb200_insert_batch(h, images, rows.data(), (int)n) != RFB200_OK) REPORT_ERROR(ERR_UNCLASSIFIED, rfb200_last_error(h));
        Image<float> Vout;
        Vout().initZeros(Xdim, Xdim, Xdim);
        if (rfb200_finalize(h, MULTIDIM_ARRAY(Vout())) != RFB200_OK) REPORT_ERROR(ERR_UNCLASSIFIED, rfb200_last_error(h));
        Vout.write(fn_out);
        rfb200_destroy(h);
        rfb200_host_free(images);
    }
};

#endif  // WITH_XMIPPCORE
