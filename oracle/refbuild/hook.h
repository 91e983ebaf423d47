/* TEST INFRASTRUCTURE ONLY (oracle/).  Name->pointer lookup for the file-static
 * arrays of a reference solver.  Each wrap_*.c does `#include "<solver>.c"`
 * (resolved through -I/root/reference, i.e. the reference source is compiled
 * where it lies, never copied) and then lists the statics it wants visible. */
#ifndef ORACLE_REF_HOOK_H
#define ORACLE_REF_HOOK_H
#include <string.h>
#define HOOK_BEGIN(fn) void *fn(const char *name) {
#define HOOK(sym) if (strcmp(name, #sym) == 0) return (void *)(sym);
#define HOOK_END return (void *)0; }
#endif
