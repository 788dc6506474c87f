// Thread-local error text behind mof_last_error_string() (include/mof_b200.h).
#ifndef MOF_ERROR_H
#define MOF_ERROR_H
#ifdef __cplusplus
extern "C" {
#endif
// Stores a printf-formatted message for the calling thread and returns `code`.
int mof_set_error(int code, const char* fmt, ...);
#ifdef __cplusplus
}
#endif
#endif
