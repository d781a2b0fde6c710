/* TEST INFRASTRUCTURE ONLY — minimal hand-declared OpenGL 3.3 core / GLX loader for oracle/gl/ref_gl_harness.cpp (the image has no
 * GL headers). Entry points are fetched with glXGetProcAddressARB from the Mesa llvmpipe libGL that ships with Nsight Compute. */
#pragma once
#include <stddef.h>
#include <stdint.h>

typedef unsigned int GLenum, GLuint, GLbitfield;
typedef int GLint, GLsizei;
typedef unsigned char GLboolean, GLubyte;
typedef float GLfloat;
typedef char GLchar;
typedef ptrdiff_t GLsizeiptr, GLintptr;

#define GL_FALSE 0
#define GL_TRUE 1
#define GL_POINTS 0x0000
#define GL_DEPTH_BUFFER_BIT 0x00000100
#define GL_COLOR_BUFFER_BIT 0x00004000
#define GL_LESS 0x0201
#define GL_DEPTH_TEST 0x0B71
#define GL_UNPACK_ALIGNMENT 0x0CF5
#define GL_PACK_ALIGNMENT 0x0D05
#define GL_TEXTURE_2D 0x0DE1
#define GL_UNSIGNED_BYTE 0x1401
#define GL_UNSIGNED_SHORT 0x1403
#define GL_UNSIGNED_INT 0x1405
#define GL_FLOAT 0x1406
#define GL_RED 0x1903
#define GL_RGB 0x1907
#define GL_RGBA 0x1908
#define GL_NEAREST 0x2600
#define GL_LINEAR 0x2601
#define GL_TEXTURE_MAG_FILTER 0x2800
#define GL_TEXTURE_MIN_FILTER 0x2801
#define GL_TEXTURE_WRAP_S 0x2802
#define GL_TEXTURE_WRAP_T 0x2803
#define GL_CLAMP_TO_EDGE 0x812F
#define GL_RGBA8 0x8058
#define GL_DEPTH_COMPONENT24 0x81A6
#define GL_R32F 0x822E
#define GL_R16UI 0x8234
#define GL_R32UI 0x8236
#define GL_RGBA32F 0x8814
#define GL_RED_INTEGER 0x8D94
#define GL_TEXTURE0 0x84C0
#define GL_PROGRAM_POINT_SIZE 0x8642
#define GL_POINT_SPRITE 0x8861
#define GL_ARRAY_BUFFER 0x8892
#define GL_STREAM_DRAW 0x88E0
#define GL_STREAM_COPY 0x88E2
#define GL_STATIC_DRAW 0x88E4
#define GL_FRAGMENT_SHADER 0x8B30
#define GL_VERTEX_SHADER 0x8B31
#define GL_GEOMETRY_SHADER 0x8DD9
#define GL_COMPILE_STATUS 0x8B81
#define GL_LINK_STATUS 0x8B82
#define GL_INTERLEAVED_ATTRIBS 0x8C8C
#define GL_TRANSFORM_FEEDBACK_BUFFER 0x8C8E
#define GL_TRANSFORM_FEEDBACK_PRIMITIVES_WRITTEN 0x8C88
#define GL_RASTERIZER_DISCARD 0x8C89
#define GL_QUERY_RESULT 0x8866
#define GL_TRANSFORM_FEEDBACK 0x8E22
#define GL_FRAMEBUFFER 0x8D40
#define GL_RENDERBUFFER 0x8D41
#define GL_COLOR_ATTACHMENT0 0x8CE0
#define GL_DEPTH_ATTACHMENT 0x8D00
#define GL_FRAMEBUFFER_COMPLETE 0x8CD5
#define GL_POINT_SIZE_RANGE 0x0B12
#define GL_ALIASED_POINT_SIZE_RANGE 0x846D
#define GL_MAX_TEXTURE_SIZE 0x0D33

#define EFGL_FUNCS(X)                                                                                                              \
  X(void, glGenTextures, (GLsizei, GLuint*))                                                                                       \
  X(void, glDeleteTextures, (GLsizei, const GLuint*))                                                                              \
  X(void, glBindTexture, (GLenum, GLuint))                                                                                         \
  X(void, glTexImage2D, (GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum, const void*))                              \
  X(void, glTexSubImage2D, (GLenum, GLint, GLint, GLint, GLsizei, GLsizei, GLenum, GLenum, const void*))                           \
  X(void, glTexParameteri, (GLenum, GLenum, GLint))                                                                                \
  X(void, glGetTexImage, (GLenum, GLint, GLenum, GLenum, void*))                                                                   \
  X(void, glActiveTexture, (GLenum))                                                                                               \
  X(void, glGenFramebuffers, (GLsizei, GLuint*))                                                                                   \
  X(void, glBindFramebuffer, (GLenum, GLuint))                                                                                     \
  X(void, glFramebufferTexture2D, (GLenum, GLenum, GLenum, GLuint, GLint))                                                         \
  X(void, glGenRenderbuffers, (GLsizei, GLuint*))                                                                                  \
  X(void, glBindRenderbuffer, (GLenum, GLuint))                                                                                    \
  X(void, glRenderbufferStorage, (GLenum, GLenum, GLsizei, GLsizei))                                                               \
  X(void, glFramebufferRenderbuffer, (GLenum, GLenum, GLenum, GLuint))                                                             \
  X(void, glDrawBuffers, (GLsizei, const GLenum*))                                                                                 \
  X(GLenum, glCheckFramebufferStatus, (GLenum))                                                                                    \
  X(void, glViewport, (GLint, GLint, GLsizei, GLsizei))                                                                            \
  X(void, glClearColor, (GLfloat, GLfloat, GLfloat, GLfloat))                                                                      \
  X(void, glClear, (GLbitfield))                                                                                                   \
  X(void, glEnable, (GLenum))                                                                                                      \
  X(void, glDisable, (GLenum))                                                                                                     \
  X(void, glDepthFunc, (GLenum))                                                                                                   \
  X(void, glGenBuffers, (GLsizei, GLuint*))                                                                                        \
  X(void, glDeleteBuffers, (GLsizei, const GLuint*))                                                                               \
  X(void, glBindBuffer, (GLenum, GLuint))                                                                                          \
  X(void, glBufferData, (GLenum, GLsizeiptr, const void*, GLenum))                                                                 \
  X(void, glBufferSubData, (GLenum, GLintptr, GLsizeiptr, const void*))                                                            \
  X(void, glGetBufferSubData, (GLenum, GLintptr, GLsizeiptr, void*))                                                               \
  X(void, glGenVertexArrays, (GLsizei, GLuint*))                                                                                   \
  X(void, glBindVertexArray, (GLuint))                                                                                             \
  X(void, glEnableVertexAttribArray, (GLuint))                                                                                     \
  X(void, glDisableVertexAttribArray, (GLuint))                                                                                    \
  X(void, glVertexAttribPointer, (GLuint, GLint, GLenum, GLboolean, GLsizei, const void*))                                         \
  X(GLuint, glCreateShader, (GLenum))                                                                                              \
  X(void, glShaderSource, (GLuint, GLsizei, const GLchar* const*, const GLint*))                                                   \
  X(void, glCompileShader, (GLuint))                                                                                               \
  X(void, glGetShaderiv, (GLuint, GLenum, GLint*))                                                                                 \
  X(void, glGetShaderInfoLog, (GLuint, GLsizei, GLsizei*, GLchar*))                                                                \
  X(GLuint, glCreateProgram, (void))                                                                                               \
  X(void, glAttachShader, (GLuint, GLuint))                                                                                        \
  X(void, glLinkProgram, (GLuint))                                                                                                 \
  X(void, glGetProgramiv, (GLuint, GLenum, GLint*))                                                                                \
  X(void, glGetProgramInfoLog, (GLuint, GLsizei, GLsizei*, GLchar*))                                                               \
  X(void, glUseProgram, (GLuint))                                                                                                  \
  X(GLint, glGetUniformLocation, (GLuint, const GLchar*))                                                                          \
  X(void, glUniform1i, (GLint, GLint))                                                                                             \
  X(void, glUniform1f, (GLint, GLfloat))                                                                                           \
  X(void, glUniform4f, (GLint, GLfloat, GLfloat, GLfloat, GLfloat))                                                                \
  X(void, glUniformMatrix4fv, (GLint, GLsizei, GLboolean, const GLfloat*))                                                         \
  X(void, glTransformFeedbackVaryings, (GLuint, GLsizei, const GLchar* const*, GLenum))                                            \
  X(void, glGenTransformFeedbacks, (GLsizei, GLuint*))                                                                             \
  X(void, glBindTransformFeedback, (GLenum, GLuint))                                                                               \
  X(void, glBindBufferBase, (GLenum, GLuint, GLuint))                                                                              \
  X(void, glBeginTransformFeedback, (GLenum))                                                                                      \
  X(void, glEndTransformFeedback, (void))                                                                                          \
  X(void, glDrawTransformFeedback, (GLenum, GLuint))                                                                               \
  X(void, glGenQueries, (GLsizei, GLuint*))                                                                                        \
  X(void, glBeginQuery, (GLenum, GLuint))                                                                                          \
  X(void, glEndQuery, (GLenum))                                                                                                    \
  X(void, glGetQueryObjectuiv, (GLuint, GLenum, GLuint*))                                                                          \
  X(void, glDrawArrays, (GLenum, GLint, GLsizei))                                                                                  \
  X(void, glReadPixels, (GLint, GLint, GLsizei, GLsizei, GLenum, GLenum, void*))                                                   \
  X(void, glFinish, (void))                                                                                                        \
  X(void, glPixelStorei, (GLenum, GLint))                                                                                          \
  X(void, glPointSize, (GLfloat))                                                                                                  \
  X(void, glGetFloatv, (GLenum, GLfloat*))                                                                                         \
  X(void, glGetIntegerv, (GLenum, GLint*))                                                                                         \
  X(const GLubyte*, glGetString, (GLenum))                                                                                         \
  X(GLenum, glGetError, (void))
