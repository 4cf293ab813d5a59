// capi_util.h -- exception -> status translation for the extern "C" entry points.
#pragma once
#include <string>
#include "host_util.h"

namespace cra5 {
std::string& last_error_slot();

template <class F>
int guarded(F&& f) {
  try {
    f();
    return OK;
  } catch (const Error& e) {
    last_error_slot() = e.what();
    return e.code;
  } catch (const std::exception& e) {
    last_error_slot() = e.what();
    return ERR_INTERNAL;
  } catch (...) {
    last_error_slot() = "unknown error";
    return ERR_INTERNAL;
  }
}
}  // namespace cra5
