// oracle/sdl_stub/SDL_opengl.h -- TEST INFRASTRUCTURE: empty stand-in (see SDL.h beside it).
