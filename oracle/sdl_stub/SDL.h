// oracle/sdl_stub/SDL.h -- TEST INFRASTRUCTURE.
// libSDL2 is not installed in this image.  The reference's src/main.cpp
// includes <SDL.h> for its serial "visual" path (src/main.cpp:108-227); this
// stub supplies do-nothing stand-ins for the handful of SDL symbols that file
// names, so that the UNMODIFIED main.cpp can be compiled in place (see
// oracle/Makefile target raytracer_unmodified) and its --parallel path used to
// pin oracle/ref_driver.cpp.  Nothing here draws anything.
#ifndef SKR_SDL_STUB_H
#define SKR_SDL_STUB_H
#include <cstdlib>
#define SDL_INIT_VIDEO 0x20u
#define SDL_QUIT 0x100u
struct SDL_Window;
struct SDL_Renderer;
struct SDL_Event
{
	unsigned type;
};
static inline int SDL_Init(unsigned) { return 0; }
static inline SDL_Window *SDL_CreateWindow(const char *, int, int, int, int, unsigned) { return nullptr; }
static inline SDL_Renderer *SDL_CreateRenderer(SDL_Window *, int, unsigned) { return nullptr; }
static inline int SDL_SetRenderDrawColor(SDL_Renderer *, unsigned char, unsigned char, unsigned char, unsigned char) { return 0; }
static inline int SDL_RenderClear(SDL_Renderer *) { return 0; }
static inline void SDL_RenderPresent(SDL_Renderer *) {}
static inline int SDL_RenderDrawPoint(SDL_Renderer *, int, int) { return 0; }
// The visual path spins forever after writing its PPM (src/main.cpp:215-223)
// unless it sees SDL_QUIT; SKR_SDL_STUB_QUIT=1 makes the stub deliver one so
// that the process can leave that loop (SDL_Quit below then exits).
static inline int SDL_PollEvent(SDL_Event *e)
{
	static int quit = -1;
	if(quit < 0)
	{
		const char *q = getenv("SKR_SDL_STUB_QUIT");
		quit		  = (q && q[0] == '1') ? 1 : 0;
	}
	(void) e;
	return 0;
}
static inline void SDL_DestroyRenderer(SDL_Renderer *) {}
static inline void SDL_DestroyWindow(SDL_Window *) {}
static inline void SDL_Quit() {}
#endif
