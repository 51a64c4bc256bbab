/* stand-in for <X11/Xlib.h>: nothing from it is used by the translation units built with this shim */
#pragma once
