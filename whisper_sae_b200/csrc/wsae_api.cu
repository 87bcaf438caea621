// wsae_api.cu — ABI identification for libwsae_sm100.so (see include/wsae.h).
extern "C" int wsae_abi_version(void) { return 109; }
