// TEST STUBS: the handful of ROOT declarations macros/npsWF_gpu.C uses, so that the macro can at least be
// type-checked against include/npswf_host.hpp without ROOT (tests/test_abi_cpu.py).  Nothing here does anything.
#pragma once
#include <cstddef>
typedef double Double_t;
typedef float Float_t;
typedef int Int_t;
class TTree {
public:
    template <class T> int Branch(const char *, T *) { return 0; }
    int Fill() { return 0; }
};
class TFile {};
class TTreeReader {
public:
    explicit TTreeReader(TTree *) {}
    bool Next() { return false; }
};
template <class T> class TTreeReaderValue {
public:
    TTreeReaderValue(TTreeReader &, const char *) {}
    T &operator*() { return v_; }
private:
    T v_{};
};
template <class T> class TTreeReaderArray {
public:
    TTreeReaderArray(TTreeReader &, const char *) {}
    size_t GetSize() const { return 0; }
    T &operator[](size_t) { return v_; }
private:
    T v_{};
};
