/* Build shim (test infrastructure): the reference includes <zlog.h>, which is not on this image.
 * Logging calls compile to nothing. */
#ifndef PF_SHIM_ZLOG_H
#define PF_SHIM_ZLOG_H
typedef struct zlog_category_s zlog_category_t;
#define dzlog_init(a, b) (0)
#define dzlog_debug(...) ((void)0)
#define dzlog_info(...) ((void)0)
#define dzlog_warn(...) ((void)0)
#define dzlog_error(...) ((void)0)
#define zlog_debug(...) ((void)0)
#define zlog_info(...) ((void)0)
#define zlog_warn(...) ((void)0)
#define zlog_error(...) ((void)0)
#define zlog_level(...) ((void)0)
#define zlog_fini() ((void)0)
#endif
