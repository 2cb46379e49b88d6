// base_options.h — one-letter plan flags (reference src/methods/base_options.h).
//   BLAS_Op  "N" / "T"   leave the operand as stored / materialise its transpose first
//   Pad_Op   "N" / "P"   leave ld alone / copy into a buffer with ld rounded up to 32 elements
//   Bool_Op  "F" / "T"
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

namespace rtat {

// A two-valued flag whose values have one-letter names.
template <class Derived, char Off, char On>
class Letter_Flag {
 public:
  bool set = false;
  Letter_Flag() = default;
  explicit Letter_Flag(bool set) : set(set) {}
  static Derived parse(const std::string &c, const char *what) {
    if (c.size() == 1 && c[0] == Off) return Derived(false);
    if (c.size() == 1 && c[0] == On) return Derived(true);
    throw std::runtime_error(std::string("Invalid ") + what + " string " + c);
  }
  operator std::string() const { return std::string(1, set ? On : Off); }
  Derived operator!() const { return Derived(!set); }
  static std::vector<Derived> enumerate() { return {Derived(false), Derived(true)}; }
};

class BLAS_Op : public Letter_Flag<BLAS_Op, 'N', 'T'> {
 public:
  enum _BLAS_Op { NOTRANS, TRANS };
  BLAS_Op() = default;
  BLAS_Op(_BLAS_Op o) : Letter_Flag(o == TRANS) {}
  explicit BLAS_Op(bool t) : Letter_Flag(t) {}
  BLAS_Op(const std::string &c) : Letter_Flag(parse(c, "BLAS_Op")) {}
  bool operator==(_BLAS_Op o) const { return set == (o == TRANS); }
};

class Pad_Op : public Letter_Flag<Pad_Op, 'N', 'P'> {
 public:
  enum _Pad_Op { NOPAD, PAD };
  Pad_Op() = default;
  Pad_Op(_Pad_Op o) : Letter_Flag(o == PAD) {}
  explicit Pad_Op(bool p) : Letter_Flag(p) {}
  Pad_Op(const std::string &c) : Letter_Flag(parse(c, "Pad_Op")) {}
  bool operator==(_Pad_Op o) const { return set == (o == PAD); }
};

class Bool_Op : public Letter_Flag<Bool_Op, 'F', 'T'> {
 public:
  Bool_Op() = default;
  Bool_Op(bool b) : Letter_Flag(b) {}
  Bool_Op(const std::string &c) : Letter_Flag(parse(c, "bool op")) {}
  operator bool() const { return set; }
};

}  // namespace rtat
