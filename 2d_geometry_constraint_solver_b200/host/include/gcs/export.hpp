// Symbol visibility of the host mirror (same macro name as the reference,
// includes/gcs/export.hpp:12-14, so client code compiles unchanged).
#pragma once
#if defined(__GNUC__) && __GNUC__ >= 4
#define GCS_API __attribute__((visibility("default")))
#define GCS_NO_EXPORT __attribute__((visibility("hidden")))
#else
#define GCS_API
#define GCS_NO_EXPORT
#endif
