// json_encoding.h — wire format of keys and plans (reference src/methods/json_encoding.h:14-108).
//   GEMM_Key          {"transA","transB","m","n","k"}
//   GEMM_Options      {"transA","transB","transC"}
//   GEMM_Options_Pad  {"transA","transB","transC","padA","padB","padC"}
//   SYRK_Key          {"uplo","trans","n","k"}          (:110-138)
//   SYRK_Options      {"transA","transC"}               (:140-162)
//   TRSM_Key          {"side","uplo","trans","diag","m","n"}   (:164-199)
//   TRSM_Options      {"swap_side","transA"}            (:201-223)
#pragma once
#include "gemm.h"
#include "json.h"
#include "syrk.h"
#include "trsm.h"

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

inline Json to_json(const SYRK_Key &key) {
  Json j = Json::object();
  j["uplo"] = std::string(key.uplo);
  j["trans"] = std::string(key.trans);
  j["n"] = key.n;
  j["k"] = key.k;
  return j;
}
template <>
inline SYRK_Key from_json<SYRK_Key>(const Json &j) {
  return SYRK_Key(BLAS_Fill_Mode(j["uplo"].get<std::string>()), BLAS_Operation(j["trans"].get<std::string>()), j["n"].get<int>(),
                  j["k"].get<int>());
}

inline Json to_json(const SYRK_Options &o) {
  Json j = Json::object();
  j["transA"] = std::string(o.transpose_A);
  j["transC"] = std::string(o.transpose_C);
  return j;
}
template <>
inline SYRK_Options from_json<SYRK_Options>(const Json &j) {
  return SYRK_Options(Bool_Op(j["transA"].get<std::string>()), Bool_Op(j["transC"].get<std::string>()));
}

inline Json to_json(const TRSM_Key &key) {
  Json j = Json::object();
  j["side"] = std::string(key.side);
  j["uplo"] = std::string(key.uplo);
  j["trans"] = std::string(key.trans);
  j["diag"] = std::string(key.diag);
  j["m"] = key.m;
  j["n"] = key.n;
  return j;
}
template <>
inline TRSM_Key from_json<TRSM_Key>(const Json &j) {
  return TRSM_Key(BLAS_Side(j["side"].get<std::string>()), BLAS_Fill_Mode(j["uplo"].get<std::string>()),
                  BLAS_Operation(j["trans"].get<std::string>()), BLAS_Diag(j["diag"].get<std::string>()), j["m"].get<int>(), j["n"].get<int>());
}

inline Json to_json(const TRSM_Options &o) {
  Json j = Json::object();
  j["swap_side"] = std::string(o.swap_side);
  j["transA"] = std::string(o.transpose_A);
  return j;
}
template <>
inline TRSM_Options from_json<TRSM_Options>(const Json &j) {
  return TRSM_Options(Bool_Op(j["swap_side"].get<std::string>()), Bool_Op(j["transA"].get<std::string>()));
}

}  // namespace rtat
