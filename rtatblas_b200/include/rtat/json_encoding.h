// json_encoding.h — wire format of keys and plans (reference src/methods/json_encoding.h:14-108).
//   GEMM_Key          {"transA","transB","m","n","k"}
//   GEMM_Options      {"transA","transB","transC"}
//   GEMM_Options_Pad  {"transA","transB","transC","padA","padB","padC"}
#pragma once
#include "gemm.h"
#include "json.h"

namespace rtat {

template <typename T>
T from_json(const Json &);

inline Json to_json(const GEMM_Key &key) {
  Json j = Json::object();
  j["transA"] = std::string(key.transa);
  j["transB"] = std::string(key.transb);
  j["m"] = key.m;
  j["n"] = key.n;
  j["k"] = key.k;
  return j;
}
template <>
inline GEMM_Key from_json<GEMM_Key>(const Json &j) {
  return GEMM_Key(BLAS_Operation(j["transA"].get<std::string>()), BLAS_Operation(j["transB"].get<std::string>()), j["m"].get<int>(),
                  j["k"].get<int>(), j["n"].get<int>());
}

inline Json to_json(const GEMM_Options &o) {
  Json j = Json::object();
  j["transA"] = std::string(o.transa);
  j["transB"] = std::string(o.transb);
  j["transC"] = std::string(o.transc);
  return j;
}
template <>
inline GEMM_Options from_json<GEMM_Options>(const Json &j) {
  return GEMM_Options(BLAS_Op(j["transA"].get<std::string>()), BLAS_Op(j["transB"].get<std::string>()), BLAS_Op(j["transC"].get<std::string>()));
}

inline Json to_json(const GEMM_Options_Pad &o) {
  Json j = Json::object();
  j["transA"] = std::string(o.transa);
  j["transB"] = std::string(o.transb);
  j["transC"] = std::string(o.transc);
  j["padA"] = std::string(o.pada);
  j["padB"] = std::string(o.padb);
  j["padC"] = std::string(o.padc);
  return j;
}
template <>
inline GEMM_Options_Pad from_json<GEMM_Options_Pad>(const Json &j) {
  return GEMM_Options_Pad(BLAS_Op(j["transA"].get<std::string>()), Pad_Op(j["padA"].get<std::string>()), BLAS_Op(j["transB"].get<std::string>()),
                          Pad_Op(j["padB"].get<std::string>()), BLAS_Op(j["transC"].get<std::string>()), Pad_Op(j["padC"].get<std::string>()));
}

}  // namespace rtat
