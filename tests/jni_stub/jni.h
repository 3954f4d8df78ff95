/* TEST INFRASTRUCTURE: the handful of JNI declarations java/sicelore_gpu_jni.c uses, so that the glue can be compile- and
 * link-checked in an image without a JDK (tests/test_abi.py).  Never shipped; a real build uses the JDK's jni.h. */
#ifndef SLR_JNI_STUB_H
#define SLR_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef uint8_t jboolean; typedef jint jsize;
typedef void *jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jarray;
typedef jarray jlongArray; typedef jarray jintArray;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    void *(*GetDirectBufferAddress)(JNIEnv *, jobject);
    jlong (*GetDirectBufferCapacity)(JNIEnv *, jobject);
    jsize (*GetArrayLength)(JNIEnv *, jarray);
    jlong *(*GetLongArrayElements)(JNIEnv *, jlongArray, jboolean *);
    jint *(*GetIntArrayElements)(JNIEnv *, jintArray, jboolean *);
    void (*ReleaseLongArrayElements)(JNIEnv *, jlongArray, jlong *, jint);
    void (*ReleaseIntArrayElements)(JNIEnv *, jintArray, jint *, jint);
    jstring (*NewStringUTF)(JNIEnv *, const char *);
};
#endif
