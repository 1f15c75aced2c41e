/* Replacement for the reference's tests/hook.c when its bench harness is
 * linked against a library that does not expose planner internals (ours).
 * Same idea as tools/fftw-wisdom.c:35-36 in the reference. */
void install_hook(void) {}
void uninstall_hook(void) {}
