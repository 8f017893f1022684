/* Empty stand-in for <winbase.h> (raytracer3.0.06's raytracer.cpp includes it but uses nothing from it).
 * TEST INFRASTRUCTURE ONLY (oracle/Makefile, oracle/_ref). */
