/* empty shim: ac3enc.cpp includes <windows.h> but uses nothing from it */
