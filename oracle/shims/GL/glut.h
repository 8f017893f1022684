/* No-op stand-in for <GL/glut.h>: lets the reference's displayfunc.cpp (which
 * holds ReadScene and UpdateCamera next to its GLUT callbacks) compile headless.
 * TEST INFRASTRUCTURE ONLY (used by oracle/Makefile to build oracle/_ref). */
#ifndef ORACLE_SHIM_GLUT_H
#define ORACLE_SHIM_GLUT_H
#define GLUT_BITMAP_HELVETICA_18 ((void *)0)
#define GLUT_KEY_UP 101
#define GLUT_KEY_DOWN 103
#define GLUT_KEY_LEFT 100
#define GLUT_KEY_RIGHT 102
#define GLUT_KEY_PAGE_UP 104
#define GLUT_KEY_PAGE_DOWN 105
#define GLUT_RGB 0
#define GLUT_DOUBLE 2
#define GL_BLEND 0
#define GL_SRC_ALPHA 0
#define GL_ONE_MINUS_SRC_ALPHA 0
#define GL_COLOR_BUFFER_BIT 0
#define GL_RGBA 0
#define GL_UNSIGNED_BYTE 0
#define GL_PROJECTION 0
#define GL_MODELVIEW 0
#define glutBitmapCharacter(...) ((void)0)
#define glutPostRedisplay(...) ((void)0)
#define glutSwapBuffers(...) ((void)0)
#define glutReshapeWindow(...) ((void)0)
#define glutInitWindowSize(...) ((void)0)
#define glutInitWindowPosition(...) ((void)0)
#define glutInitDisplayMode(...) ((void)0)
#define glutInit(...) ((void)0)
#define glutCreateWindow(...) ((void)0)
#define glutReshapeFunc(...) ((void)0)
#define glutKeyboardFunc(...) ((void)0)
#define glutSpecialFunc(...) ((void)0)
#define glutDisplayFunc(...) ((void)0)
#define glutIdleFunc(...) ((void)0)
#define glutMainLoop(...) ((void)0)
#define glEnable(...) ((void)0)
#define glDisable(...) ((void)0)
#define glBlendFunc(...) ((void)0)
#define glColor4f(...) ((void)0)
#define glColor3f(...) ((void)0)
#define glRecti(...) ((void)0)
#define glRasterPos2i(...) ((void)0)
#define glClear(...) ((void)0)
#define glDrawPixels(...) ((void)0)
#define glViewport(...) ((void)0)
#define glLoadIdentity(...) ((void)0)
#define glOrtho(...) ((void)0)
#define glMatrixMode(...) ((void)0)
#define glClearColor(...) ((void)0)
#define glPushMatrix(...) ((void)0)
#define glPopMatrix(...) ((void)0)
#endif
