// Library identification for libqa_b200.so.
#include "qa_b200.h"

extern "C" int qa_version(void) { return QA_ABI_VERSION; }

extern "C" const char* qa_build_info(void) {
    return "libqa_b200 abi " 
#define QA_STR2(x) #x
#define QA_STR(x) QA_STR2(x)
        QA_STR(QA_ABI_VERSION) ", sm_100a, nvcc " QA_STR(__CUDACC_VER_MAJOR__) "." QA_STR(__CUDACC_VER_MINOR__)
        ", built " __DATE__;
}

// Layout handshake for FFI bindings: sizeof() of every argument struct, so that a ctypes / cgo /
// JNI mirror can assert it agrees with this build before passing pointers.
extern "C" int qa_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(QaActionPushArgs);
        case 1: return (int)sizeof(QaTorqueArgs);
        case 2: return (int)sizeof(QaTerrain);
        case 3: return (int)sizeof(QaHeightScanArgs);
        case 4: return (int)sizeof(QaMocapTable);
        case 5: return (int)sizeof(QaMocapBlendArgs);
        case 6: return (int)sizeof(QaBbcConst);
        case 7: return (int)sizeof(QaBbcStepArgs);
        case 8: return (int)sizeof(QaCompactArgs);
        case 9: return (int)sizeof(QaGaeArgs);
        case 10: return (int)sizeof(QaGatherArgs);
        case 11: return (int)sizeof(QaClipAdamArgs);
        case 12: return (int)sizeof(QaLinearArgs);
        case 13: return (int)sizeof(QaActBwdArgs);
        case 14: return (int)sizeof(QaPpoLossArgs);
        case 15: return (int)sizeof(QaLinearBwdArgs);
        case 16: return (int)sizeof(QaHistEncArgs);
        case 17: return (int)sizeof(QaRowLossArgs);
        case 18: return (int)sizeof(QaPpoScalarsArgs);
        case 19: return (int)sizeof(QaDepthArgs);
        case 20: return (int)sizeof(QaPpoLossTscArgs);
        case 21: return (int)sizeof(QaTscConst);
        case 22: return (int)sizeof(QaTscStepArgs);
        case 23: return (int)sizeof(QaDiscInputArgs);
        case 24: return (int)sizeof(QaDiscRewardArgs);
        case 25: return (int)sizeof(QaHeadFwdArgs);
        case 26: return (int)sizeof(QaHeadBwdArgs);
        case 27: return (int)sizeof(QaPolicySampleArgs);
        case 28: return (int)sizeof(QaDiscPrepareArgs);
        case 29: return (int)sizeof(QaDiscHeadsArgs);
        case 30: return (int)sizeof(QaDiscGpArgs);
        case 31: return (int)sizeof(QaDiscRegArgs);
        case 32: return (int)sizeof(QaNormMomentsArgs);
        case 33: return (int)sizeof(QaNormMergeArgs);
        case 34: return (int)sizeof(QaPeerAllreduceArgs);
        case 35: return (int)sizeof(QaAdamChainArgs);
        default: return -1;
    }
}

extern "C" int qa_zero_async(void* dst, uint64_t bytes, void* stream) {
    if (bytes == 0) return 0;
    if (dst == nullptr) return QA_EINVAL;
    return (int)cudaMemsetAsync(dst, 0, (size_t)bytes, (cudaStream_t)stream);
}

extern "C" int qa_copy_async(void* dst, const void* src, uint64_t bytes, void* stream) {
    if (bytes == 0) return 0;
    if (dst == nullptr || src == nullptr) return QA_EINVAL;
    return (int)cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}
