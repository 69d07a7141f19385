/* Empty stand-in for the Windows-only <direct.h> that ldc.cu:25 and
 * Poiseulle.cu:19 include but never use.  Lets the reference sources compile
 * unmodified on Linux (see oracle/Makefile). */
#pragma once
